"""Helpers to compare matrices with the reference's 2-decimal printed goldens
(tests/tp_02.cc:19-27 of the reference: '%7.2f', blank when |x| < 0.01)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def matches_print(M, G):
    M = np.atleast_2d(np.asarray(M, float))
    if M.shape != (len(G), len(G[0])):
        return False
    for i, row in enumerate(G):
        for j, g in enumerate(row):
            if g is None:
                if not abs(M[i, j]) < 0.01 + 1e-9:
                    return False
            elif abs(M[i, j] - g) > 0.005 + 1e-9:
                return False
    return True
