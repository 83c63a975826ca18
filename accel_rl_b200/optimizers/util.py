"""reference: accel_rl/optimizers/util.py:8-18"""
import numpy as np


def iterate_mb_idxs(batch_size, data_length, shuffle=False):
    """Yields minibatch index tuples; the shuffle consumes the global legacy numpy stream exactly like
    the reference (np.random.shuffle of arange(data_length)); the tail is dropped."""
    if shuffle:
        indices = np.arange(data_length)
        np.random.shuffle(indices)
    for start_idx in range(0, data_length - batch_size + 1, batch_size):
        if shuffle:
            yield (indices[start_idx:start_idx + batch_size],)
        else:
            yield (start_idx, start_idx + batch_size)


def epoch_index_block(batch_size, data_length, epochs, shuffle=True):
    """All minibatch indices of `epochs` epochs as one int32 array [epochs * n_mb * batch_size], in the
    order iterate_mb_idxs would yield them (same RNG consumption)."""
    n_mb = data_length // batch_size
    out = np.empty((epochs, n_mb * batch_size), np.int32)
    for ep in range(epochs):
        indices = np.arange(data_length)
        if shuffle:
            np.random.shuffle(indices)
        out[ep] = indices[:n_mb * batch_size]
    return out.reshape(-1), n_mb
