"""Mirror of heal_swin/models_torch/hp_shifting.py: the four shift strategies.

The permutation tables and mask group ids are built by the C-ABI index layer
(csrc/hs_index.cpp; milliseconds instead of the reference's per-window Python loops, and without
healpy).  ``shift`` / ``shift_back`` are provided for API compatibility and run the row-gather
kernel; inside ``SwinTransformerBlock`` the permutation is folded into the attention kernel's
loads and stores instead, so no shifted copy of the activations is ever materialised.
"""
import torch

from .. import _lib, hp_index
from ..ops import GatherRows


def get_attn_mask_from_mask(mask, window_size):
    """(N,) group ids -> (nW, ws, ws) additive mask in {0, -100}   [hp_shifting.py:10-28]"""
    return hp_index.attn_mask_from_groups(mask, window_size)


class _TableShift:
    """Common machinery: int64 tables on the host (reference attributes ``shift_idcs`` /
    ``back_shift_idcs``), lazily mirrored as int32 on whatever device the activations live on."""

    shift_idcs = None
    back_shift_idcs = None
    groups = None  # int8 (N,), ids of the shifted pixels

    def __init__(self):
        self._dev = {}

    def device_tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (
                self.shift_idcs.to(device=device, dtype=torch.int32),
                self.back_shift_idcs.to(device=device, dtype=torch.int32),
                self.groups.to(device=device, dtype=torch.uint8),
            )
        return self._dev[key]

    def shift(self, x):
        fwd, back, _ = self.device_tables(x.device)
        return GatherRows.apply(x, fwd, back)

    def shift_back(self, x):
        fwd, back, _ = self.device_tables(x.device)
        return GatherRows.apply(x, back, fwd)


class NoShift:
    """hp_shifting.py:31-39"""

    shift_idcs = None
    back_shift_idcs = None
    groups = None

    def get_mask(self):
        return None

    def shift(self, x):
        return x

    def shift_back(self, x):
        return x


class NestRollShift(_TableShift):
    """Roll by ``shift_size`` along the nested index   [hp_shifting.py:42-73]"""

    def __init__(self, shift_size, input_resolution, window_size):
        super().__init__()
        self.shift_size = shift_size
        self.input_resolution = input_resolution
        self.window_size = window_size
        self.shift_idcs, self.back_shift_idcs, self.groups = _roll_tables(shift_size, input_resolution, window_size)

    def get_mask(self):
        return get_attn_mask_from_mask(self.groups, self.window_size)


def _roll_tables(shift_size, N, window_size):
    # closed form of hp_shifting.py:48-73 (torch.roll(x, -s)[p] == x[(p + s) % N]); tiny, host side
    p = torch.arange(N, dtype=torch.int64)
    fwd = (p + shift_size) % N
    back = (p - shift_size) % N
    groups = torch.zeros(N, dtype=torch.int8)
    groups[max(N - window_size, 0): N - shift_size] = 1
    groups[N - shift_size:] = 2
    return fwd, back, groups


class NestGridShift(_TableShift):
    """Half-window shift along both HEALPix grid directions   [hp_shifting.py:76-306]"""

    def __init__(self, nside, base_pix, window_size):
        super().__init__()
        self.nside, self.base_pix, self.ws = nside, base_pix, window_size
        self.npix = base_pix * nside**2
        self.shift_idcs, self.back_shift_idcs, self.groups = hp_index.shift_tables(
            _lib.SHIFT_NEST_GRID, nside, base_pix, window_size, 0)

    def get_mask(self, get_attn_mask=True):
        if get_attn_mask:
            return get_attn_mask_from_mask(self.groups, self.ws)
        return self.groups.to(torch.float32)


class RingShift(_TableShift):
    """Roll in RING order, mapped back to NESTED   [hp_shifting.py:309-404]"""

    def __init__(self, nside, base_pix, window_size, shift_size):
        super().__init__()
        self.nside, self.base_pix, self.ws, self.shift_size = nside, base_pix, window_size, shift_size
        self.npix = base_pix * nside**2
        self.shift_idcs, self.back_shift_idcs, self.groups = hp_index.shift_tables(
            _lib.SHIFT_RING, nside, base_pix, window_size, shift_size)
        self.mask = self.groups.to(torch.int64)

    def get_mask(self, get_attn_mask=True):
        if get_attn_mask:
            # the reference keeps its ring mask ids as int64 (hp_shifting.py:380), so the derived
            # attention mask buffer is int64 {0, -100} too; keep the dtype for checkpoint parity
            return get_attn_mask_from_mask(self.groups, self.ws).to(torch.int64)
        return self.mask
