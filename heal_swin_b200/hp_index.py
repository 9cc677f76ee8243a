"""Host-side HEALPix index layer (thin wrappers over the C-ABI index functions).

All tables are built by libhealswin_b200 (csrc/hs_index.cpp); results are returned as torch
CPU tensors with the dtypes the reference uses (int64 indices, fp32 masks).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib

_vp = C.c_void_p


def _np_ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def nest_win_idcs(window_size: int) -> torch.Tensor:
    """hp_windowing.get_nest_win_idcs (hp_windowing.py:43-62)."""
    S = int(window_size**0.5)
    out = np.zeros((S, S), dtype=np.int64)
    _lib.check(_lib.lib.hs_nest_win_idcs(int(window_size), _np_ptr(out)))
    return torch.from_numpy(out)


def rel_pos_index(window_size: int) -> torch.Tensor:
    """The relative_position_index buffer of WindowAttention (swin_hp_transformer.py:98-114)."""
    out = np.zeros((window_size, window_size), dtype=np.int64)
    _lib.check(_lib.lib.hs_rel_pos_index(int(window_size), _np_ptr(out)))
    return torch.from_numpy(out)


def nest2ring(nside: int, ipix) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(ipix, dtype=np.int64))
    out = np.empty_like(a)
    _lib.check(_lib.lib.hs_nest2ring(int(nside), _np_ptr(a), _np_ptr(out), a.size))
    return out


def ring2nest(nside: int, ipix) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(ipix, dtype=np.int64))
    out = np.empty_like(a)
    _lib.check(_lib.lib.hs_ring2nest(int(nside), _np_ptr(a), _np_ptr(out), a.size))
    return out


def shift_tables(strategy: int, nside: int, base_pix: int, window_size: int, shift_size: int,
                 want_idcs: bool = True, want_groups: bool = True):
    """(shift_idcs int64 (N,), back_idcs int64 (N,), groups int8 (N,)) for one block."""
    N = base_pix * nside * nside
    idcs = np.empty(N, dtype=np.int64) if want_idcs else None
    back = np.empty(N, dtype=np.int64) if want_idcs else None
    grp = np.empty(N, dtype=np.int8) if want_groups else None
    _lib.check(_lib.lib.hs_shift_tables(int(strategy), int(nside), int(base_pix), int(window_size),
                                        int(shift_size), _np_ptr(idcs), _np_ptr(back), _np_ptr(grp)))
    cv = lambda a: torch.from_numpy(a) if a is not None else None  # noqa: E731
    return cv(idcs), cv(back), cv(grp)


def attn_mask_from_groups(groups: torch.Tensor, window_size: int) -> torch.Tensor:
    """hp_shifting.get_attn_mask_from_mask (hp_shifting.py:10-28)."""
    g = np.ascontiguousarray(groups.detach().cpu().numpy().astype(np.int8))
    N = g.shape[0]
    out = np.empty((N // window_size, window_size, window_size), dtype=np.float32)
    _lib.check(_lib.lib.hs_attn_mask_from_groups(_np_ptr(g), N, int(window_size), _np_ptr(out)))
    return torch.from_numpy(out)


def nside_of(input_resolution: int, base_pix: int) -> int:
    """swin_hp_transformer.py:271-274."""
    nside = math.sqrt(input_resolution // base_pix)
    assert nside % 1 == 0, "nside has to be an integer in every layer"
    return int(nside)
