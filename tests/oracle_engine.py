"""Test-only stand-in for ``HostEngine`` that runs the C oracle.

Lets the CPU suite exercise the host-side logic of the drop-in classes (ingest,
metadata, side effects) against the real reference without a GPU.  Never used
by the product.
"""

from __future__ import annotations

import numpy as np

from oracle.oracle import oracle_tick
from townlet_b200 import capi


class OracleHostEngine:
    def __init__(self, obs, rew, device: int = 0) -> None:
        self.obs, self.rew = obs, rew

    def alloc_outputs(self, batch, phases, n_ticks=1, *, summary=True, comps=True, pin=False):
        return {}

    def tick(self, batch, out, phases, n_ticks=1) -> None:
        res = oracle_tick(batch, self.obs, self.rew, phases, n_ticks)
        out.clear()
        out.update(res)

    def close(self) -> None:
        pass
