"""Mirror of the hot-path part of reference core/utils.py."""
import numpy as np


def get_anchors(anchors_file):
    """reference core/utils.py:31-37: ``loadtxt(delimiter=',')`` then ``reshape(-1, 3, 2)``; row block 0 belongs to the
    coarsest grid.  Values are fractions of the image side."""
    nanchors_per_scale = 3
    anchor_entry_size = 2
    anchors_table = np.loadtxt(anchors_file, dtype=float, delimiter=',')
    anchors_table = anchors_table.reshape(-1, nanchors_per_scale, anchor_entry_size)
    return anchors_table


def count_file_lines(filename):
    """reference core/utils.py:40-43 (``nclasses = len(lines)`` -- 38 for datasets/pets_breed.names)."""
    with open(filename, 'r') as fp:
        nlines = len(fp.readlines())
    return nlines
