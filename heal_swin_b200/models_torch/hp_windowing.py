"""Mirror of heal_swin/models_torch/hp_windowing.py.

In HEALPix NESTED order a window is a contiguous run of ``window_size`` tokens, so partition and
reverse are pure views (hp_windowing.py:18-21, 37-40); the index table comes from the C-ABI
index layer.
"""
from .. import hp_index


def _check_pow2(window_size):
    # hp_windowing.py:16,35
    assert window_size > 0 and (window_size & (window_size - 1)) == 0


def window_partition(x, window_size):
    """(B, N, C) -> (num_windows*B, window_size, C)   [hp_windowing.py:6-21]"""
    _check_pow2(window_size)
    B, N, C = x.shape
    return x.contiguous().view(B * (N // window_size), window_size, C)


def window_reverse(windows, window_size, N):
    """(num_windows*B, window_size, C) -> (B, N, C)   [hp_windowing.py:24-40]"""
    _check_pow2(window_size)
    B = int(windows.shape[0] / (N // window_size))
    return windows.contiguous().view(B, N, -1)


def get_nest_win_idcs(window_size):
    """sqrt(ws) x sqrt(ws) int64 tensor of nested indices   [hp_windowing.py:43-62]"""
    return hp_index.nest_win_idcs(window_size)
