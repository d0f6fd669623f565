"""Console helpers used by the generation CLI (reference `wavernn/utils/display.py:9-58`); no matplotlib."""
from __future__ import annotations

import sys


def progbar(i, n, size=16):
    done = (i * size) // n
    return ''.join('█' if k <= done else '░' for k in range(size))


def stream(message):
    sys.stdout.write(f'\r{message}')


def simple_table(item_tuples):
    """Prints a one-row table: headings over cells, each column as wide as its longer entry."""
    cols = []
    for heading, cell in item_tuples:
        heading, cell = str(heading), str(cell)
        width = max(len(heading), len(cell))
        cols.append((heading.center(width), cell.center(width)))
    border = ''.join('+' + '-' * (len(h) + 2) for h, _ in cols) + '+'
    head = ''.join(f'| {h} ' for h, _ in cols) + '|'
    body = ''.join(f'| {c} ' for _, c in cols) + '|'
    print(border)
    print(head)
    print(border)
    print(body)
    print(border)
    print(' ')
