"""Minimal tabular logger with the call surface the runners use from rllab.misc.logger
(reference: rllab/misc/logger.py:201 record_tabular, :261 dump_tabular, :319 save_itr_params,
:439 record_tabular_misc_stat).  Writes progress.csv with the reference's column names."""
import csv
import os
import time
from contextlib import contextmanager

import numpy as np

_prefixes = []
_tabular = []
_csv_path = None
_csv_header = None
_snapshot_dir = None
_snapshot_mode = "none"
_snapshot_gap = 1
_quiet = False
last_row = {}


def configure(log_dir=None, snapshot_mode="none", quiet=False, snapshot_gap=1):
    global _csv_path, _csv_header, _snapshot_dir, _snapshot_mode, _quiet, _snapshot_gap
    if snapshot_mode not in ("all", "last", "gap", "none"):
        raise NotImplementedError("snapshot_mode must be one of all / last / gap / none")
    _quiet = quiet
    _snapshot_mode = snapshot_mode
    _snapshot_gap = max(1, int(snapshot_gap))
    _csv_header = None
    if log_dir is not None:
        os.makedirs(log_dir, exist_ok=True)
        _csv_path = os.path.join(log_dir, "progress.csv")
        _snapshot_dir = log_dir
        if os.path.exists(_csv_path):
            os.remove(_csv_path)
    else:
        _csv_path = None
        _snapshot_dir = None


def log(s):
    if not _quiet:
        print("%s | %s%s" % (time.strftime("%Y-%m-%d %H:%M:%S"), "".join(_prefixes), s), flush=True)


@contextmanager
def prefix(key):
    _prefixes.append(key)
    try:
        yield
    finally:
        _prefixes.pop()


def record_tabular(key, val):
    _tabular.append((str(key), val))


def record_tabular_misc_stat(key, values, placement="back"):
    if placement == "front":
        name = lambda s: s + key
    else:
        name = lambda s: key + s
    if len(values) > 0:
        v = np.asarray(values, dtype=np.float64)
        stats = (np.average(v), np.std(v), np.median(v), np.min(v), np.max(v))
    else:
        stats = (np.nan,) * 5
    for s, x in zip(("Average", "Std", "Median", "Min", "Max"), stats):
        record_tabular(name(s), x)


def dump_tabular(with_prefix=True):
    global _csv_header
    row = dict(_tabular)
    last_row.clear()
    last_row.update(row)
    if not _quiet:
        width = max(len(k) for k in row) if row else 0
        for k, v in _tabular:
            print("%s  %s" % (k.ljust(width), v))
        print("-" * (width + 16), flush=True)
    if _csv_path is not None:
        new = _csv_header is None
        if new:
            _csv_header = list(row.keys())
        with open(_csv_path, "a", newline="") as f:
            w = csv.DictWriter(f, fieldnames=_csv_header, extrasaction="ignore")
            if new:
                w.writeheader()
            w.writerow(row)
    del _tabular[:]


def save_itr_params(itr, params):
    """snapshot_mode in all/last/gap/none (reference: rllab/misc/logger.py:319-340)"""
    if _snapshot_dir is None or _snapshot_mode == "none":
        return
    import joblib
    if _snapshot_mode == "all":
        path = os.path.join(_snapshot_dir, "itr_%d.pkl" % itr)
    elif _snapshot_mode == "last":
        path = os.path.join(_snapshot_dir, "params.pkl")
    elif _snapshot_mode == "gap":                           # every snapshot_gap-th iteration, and the first
        if not (itr == 0 or (itr + 1) % _snapshot_gap == 0):
            return
        path = os.path.join(_snapshot_dir, "itr_%d.pkl" % itr)
    else:
        return
    joblib.dump(params, path, compress=3)
