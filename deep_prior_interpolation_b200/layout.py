"""Physical channel layouts of channels-last activations.

Every activation is stored as ``[voxels][C_p]`` fp32 with ``C_p`` a multiple of 4 (16-byte channel groups,
the TMA / float4 granularity).  The odd channel counts of the MultiRes blocks (4/8/13, 8/17/26, ...
``mulresunet.py:70-79``) are handled by *segmented* layouts: each branch of a block starts at a 16-byte
boundary, pad channels hold zeros, and ``phys2log`` maps a physical channel to the reference's logical
channel (or -1).  ``torch.cat`` (``mulresunet.py:31,89``; ``base.py:319,359``) then never has to move data:
branches are written straight into their slices.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


def pad4(c: int) -> int:
    return (c + 3) // 4 * 4


@dataclass(frozen=True)
class ChannelLayout:
    segments: Tuple[Tuple[int, int, int], ...]   # (logical_start, length, phys_start)
    C_l: int
    C_p: int

    @staticmethod
    def dense(c: int) -> "ChannelLayout":
        return ChannelLayout(((0, c, 0),), c, pad4(c))

    @staticmethod
    def concat(parts: Sequence["ChannelLayout"]) -> "ChannelLayout":
        segs: List[Tuple[int, int, int]] = []
        lo = po = 0
        for p in parts:
            for (ls, n, ps) in p.segments:
                segs.append((lo + ls, n, po + ps))
            lo += p.C_l
            po += p.C_p
        return ChannelLayout(tuple(segs), lo, po)

    def phys2log(self) -> np.ndarray:
        m = np.full(self.C_p, -1, dtype=np.int32)
        for (ls, n, ps) in self.segments:
            m[ps:ps + n] = np.arange(ls, ls + n, dtype=np.int32)
        return m

    def part_offsets(self, parts: Sequence["ChannelLayout"]) -> List[int]:
        """physical offsets of the parts this layout was concatenated from"""
        offs, po = [], 0
        for p in parts:
            offs.append(po)
            po += p.C_p
        return offs
