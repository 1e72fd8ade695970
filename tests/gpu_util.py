import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECK = {
    "c5g7": os.path.join(ROOT, "decks", "c5g7", "c5g7_2d"),
    "c5g7_3d": os.path.join(ROOT, "decks", "c5g7", "c5g7_3d_rodded"),
    "inf": os.path.join(ROOT, "decks", "urr", "inf"),
    "slab": os.path.join(ROOT, "decks", "urr", "slab"),
    "ce_pin": os.path.join(ROOT, "decks", "ce", "pincell"),
    "can": os.path.join(ROOT, "decks", "mg", "can"),
}


def random_points(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    r = rng.uniform(lo, hi, size=(n, 3))
    u = rng.normal(size=(n, 3))
    u /= np.sqrt((u * u).sum(1))[:, None]
    return r, u
