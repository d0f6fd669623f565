"""Console helpers of the generation CLI (names as in the reference `wavernn/utils/display.py`: progbar, stream,
simple_table); no matplotlib import, the plotting helpers of the reference are not part of the generation path."""
from __future__ import annotations

import sys

_FULL, _EMPTY = '█', '░'


def progbar(i, n, size=16):
    """`size` cells, the first (i*size)//n + 1 of them filled."""
    filled = min(size, (i * size) // max(n, 1) + 1)
    return _FULL * filled + _EMPTY * (size - filled)


def stream(message):
    """Overwrite the current console line."""
    sys.stdout.write('\r' + str(message))
    sys.stdout.flush()


def simple_table(item_tuples):
    """One-row table: headings above cells, each column as wide as its longer entry."""
    cols = []
    for heading, cell in item_tuples:
        heading, cell = str(heading), str(cell)
        width = max(len(heading), len(cell))
        cols.append((heading.center(width), cell.center(width)))
    rule = '+' + '+'.join('-' * (len(h) + 2) for h, _ in cols) + '+'
    for line in (rule, '| ' + ' | '.join(h for h, _ in cols) + ' |', rule, '| ' + ' | '.join(c for _, c in cols) + ' |', rule, ' '):
        print(line)
