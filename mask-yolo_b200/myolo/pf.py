"""Padded-flat (PF) activation layout used by the 3x3 tap-GEMM kernels.

A batch of n tiles of H x W pixels with C channels is stored as a 2-D row-major matrix
[n*(H+1)*(W+1), C]: every tile carries ONE zero row-block on top ((W+1) rows) and ONE zero pixel at
the left of each line.  Because the layout is flat, the left pad of line h+1 doubles as the right
pad of line h and the top pad of tile i+1 doubles as the bottom pad of tile i, so a 3x3 SAME
convolution becomes nine GEMMs whose A operand is the same matrix shifted by a constant number of
rows:  shift(dy,dx) = (dy-1)*(W+1) + (dx-1).  Guard rows (>= W+2, zero) precede and follow the
matrix for the fp32 kernel; the TMA path zero-fills out-of-range rows itself.

Invariant: pad rows are zero in every PF tensor, always (kernels never write them).
"""
from __future__ import annotations

import torch

from . import _cabi


def guard_rows(W: int) -> int:
    return W + 2


class PF:
    """Owns a zero-initialised PF buffer.  .rows is the [M, C] matrix (M = n*(H+1)*(W+1))."""

    def __init__(self, n, H, W, C, device="cuda", storage=None, split=False, dtype=torch.float32):
        """split=True allocates a second copy (the tf32 low part of a 3xTF32 operand) `lo_off` rows
        after the first, separated by zero guard rows: [guard | rows | guard | rows_lo | guard]."""
        self.n, self.H, self.W, self.C = n, H, W, C
        self.M = n * (H + 1) * (W + 1)
        g = guard_rows(W)
        copies = 2 if split else 1
        total = (copies * (self.M + g) + g) * C
        self.esize = torch.empty(0, dtype=dtype).element_size()      # 4 (fp32) or 2 (IEEE half, "h16" mask head)
        if storage is None:
            storage = torch.zeros(total, dtype=dtype, device=device)
        else:
            assert storage.numel() >= total
            storage = storage[:total]
        self.storage = storage
        self.rows = storage[g * C:(g + self.M) * C].view(self.M, C)
        self.lo_off = self.M + g if split else 0
        self.rows_lo = storage[(g + self.lo_off) * C:(g + self.lo_off + self.M) * C].view(self.M, C) if split else None

    def view(self, lo=False) -> _cabi.View:
        """C-ABI view of the valid pixels: pixel (0,0) of tile 0 sits (W+1)+1 rows in."""
        W, H, C = self.W, self.H, self.C
        off = ((W + 1) + 1) * C
        base = (self.rows_lo if lo else self.rows).data_ptr()
        return _cabi.View(base + self.esize * off, (H + 1) * (W + 1) * C, (W + 1) * C, self.n, H, W, C)

    def valid(self) -> torch.Tensor:
        """[n,H,W,C] strided torch view of the valid pixels (no copy)."""
        t = self.rows.view(self.n, self.H + 1, self.W + 1, self.C)
        return t[:, 1:, 1:, :]

    def load_dense(self, x: torch.Tensor):
        self.valid().copy_(x)
        return self

    def dense(self) -> torch.Tensor:
        return self.valid().contiguous()

    def zero_(self):
        self.storage.zero_()
        return self


def conv3x3_shifts(W: int, negate: bool = False):
    s = [(dy - 1) * (W + 1) + (dx - 1) for dy in range(3) for dx in range(3)]
    return [-v for v in s] if negate else s
