"""Checkpoint / output locations, same attribute names as the reference's `wavernn/utils/paths.py:6-34`."""
from __future__ import annotations

import os
from pathlib import Path


class Paths:
    def __init__(self, voc_id, base=None):
        # the reference anchors at the repository root (three levels above its utils/paths.py); here the root is
        # the working directory unless given, so the shipped `logs_wavernn/` tree is found next to wavernn_gen.py
        self.base = Path(base if base is not None else os.environ.get('B200TTS_BASE', os.getcwd())).expanduser().resolve()
        self.voc_checkpoints = self.base / 'logs_wavernn/checkpoints'
        self.voc_latest_weights = self.voc_checkpoints / 'latest_weights.pyt'
        self.voc_latest_optim = self.voc_checkpoints / 'latest_optim.pyt'
        self.voc_output = self.base / 'logs_wavernn/model_outputs'
        self.voc_step = self.voc_checkpoints / 'step.npy'
        self.voc_log = self.voc_checkpoints / 'log.txt'
        self.create_paths()

    def create_paths(self):
        os.makedirs(self.voc_checkpoints, exist_ok=True)
        os.makedirs(self.voc_output, exist_ok=True)

    def get_voc_named_weights(self, name):
        return self.voc_checkpoints / f'{name}_weights.pyt'

    def get_voc_named_optim(self, name):
        return self.voc_checkpoints / f'{name}_optim.pyt'
