"""Seeding with the reference's semantics (rllab/misc/ext.py:198-207): python `random`, the GLOBAL
numpy legacy stream (np.random.seed) and a separate RandomState(seed) that Lasagne's initialisers
draw from (GlorotUniform for the conv filters)."""
import random

import numpy as np

_conv_init_rng = np.random
_seed = None


def set_seed(seed):
    global _conv_init_rng, _seed
    seed %= 4294967294
    _seed = seed
    random.seed(seed)
    np.random.seed(seed)
    _conv_init_rng = np.random.RandomState(seed)


def get_seed():
    return _seed


def get_conv_init_rng():
    """the stream lasagne.random.get_rng() would return"""
    return _conv_init_rng
