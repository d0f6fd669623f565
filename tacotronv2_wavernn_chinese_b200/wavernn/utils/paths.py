"""Checkpoint / output locations of the vocoder.

Exposes the attribute names the reference's `wavernn/utils/paths.py` object has (`voc_checkpoints`,
`voc_latest_weights`, `voc_latest_optim`, `voc_output`, `voc_step`, `voc_log`, `get_voc_named_*`), because
`wavernn_gen.py` and user scripts read them.  Unlike the reference (anchored three directories above its own source
file) the tree is anchored at `base`, default: $B200TTS_BASE or the working directory, so the shipped
`logs_wavernn/checkpoints/latest_weights.pyt` is found next to wherever `wavernn_gen.py` is run.
"""
from __future__ import annotations

import os
from pathlib import Path

_LAYOUT = {                                   # attribute -> path below `base`
    'voc_checkpoints': 'logs_wavernn/checkpoints',
    'voc_output': 'logs_wavernn/model_outputs',
}
_CHECKPOINT_FILES = {                         # attribute -> file inside voc_checkpoints
    'voc_latest_weights': 'latest_weights.pyt',
    'voc_latest_optim': 'latest_optim.pyt',
    'voc_step': 'step.npy',
    'voc_log': 'log.txt',
}


class Paths:
    def __init__(self, voc_id, base=None):
        self.voc_id = voc_id
        root = base if base is not None else os.environ.get('B200TTS_BASE', os.getcwd())
        self.base = Path(root).expanduser().resolve()
        for attr, rel in _LAYOUT.items():
            setattr(self, attr, self.base / rel)
        for attr, name in _CHECKPOINT_FILES.items():
            setattr(self, attr, self.voc_checkpoints / name)
        self.create_paths()

    def create_paths(self):
        for attr in _LAYOUT:
            getattr(self, attr).mkdir(parents=True, exist_ok=True)

    def _named(self, name, kind):
        return self.voc_checkpoints / f'{name}_{kind}.pyt'

    def get_voc_named_weights(self, name):
        return self._named(name, 'weights')

    def get_voc_named_optim(self, name):
        return self._named(name, 'optim')
