"""Process-wide hyper-parameter singleton, same contract as the reference's
`wavernn/utils/__init__.py:40-92`:

    from tacotronv2_wavernn_chinese_b200.wavernn.utils import hparams as hp
    hp.configure('wavernn_hparams.py')      # exec the python file, copy its public names
    hp.bits, hp.hop_length, ...

`configure` is one-shot (RuntimeError on a second call, reference :60-61), attribute access before
it raises AttributeError (:51-53), a missing file raises FileNotFoundError (:66-67) and a non-.py
path ValueError (:68-69).  The reference's `data_parallel_workaround` (:22-36) is training-only and
not part of this package.
"""
from __future__ import annotations

import importlib.util
import re
from pathlib import Path
from typing import Union

_DUNDER = re.compile(r'^__.+__$')


class _HParams:
    def __init__(self):
        object.__setattr__(self, '_configured', False)

    def is_configured(self) -> bool:
        return self._configured

    def __getattr__(self, item):
        # only reached when normal lookup fails
        if not object.__getattribute__(self, '_configured'):
            raise AttributeError('HParams not configured yet. Call self.configure()')
        raise AttributeError(f'hparams has no attribute {item!r}')

    def configure(self, path: Union[str, Path]):
        if self._configured:
            raise RuntimeError('Cannot reconfigure hparams!')
        path = Path(path).expanduser()
        if not path.exists():
            raise FileNotFoundError(f'Could not find hparams file {path}')
        if path.suffix != '.py':
            raise ValueError('`path` must be a python file')
        spec = importlib.util.spec_from_file_location('hparams', path)
        if spec is None:
            raise ValueError(f'could not load module from "{path}"')
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for name, value in vars(mod).items():
            if _DUNDER.match(name):
                continue
            if name in self.__dict__:
                raise AttributeError(f'module at `path` cannot contain attribute {name} as it '
                                     'overwrites an attribute of the same name in utils.hparams')
            object.__setattr__(self, name, value)
        object.__setattr__(self, '_configured', True)


hparams = _HParams()
