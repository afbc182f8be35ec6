"""Host-side blake2s tables (strings never reach the device).

* ``last_action_hash`` — src/townlet/world/observations/encoders/features.py:194-200
* ``id_embedding_words`` — src/townlet/world/observations/encoders/social.py:236-242
"""

from __future__ import annotations

import hashlib
from functools import lru_cache

import numpy as np


@lru_cache(maxsize=4096)
def last_action_hash(last_action_id: str) -> float:
    """float64 value the reference assigns to ``last_action_id_hash`` (cast to f32 on store)."""
    if not isinstance(last_action_id, str) or not last_action_id:
        return 0.0
    digest = hashlib.blake2s(last_action_id.encode("utf-8")).digest()
    return int.from_bytes(digest[:4], "little") / float(2**32)


@lru_cache(maxsize=65536)
def _digest_words(agent_id: str) -> tuple[int, int, int, int]:
    digest = hashlib.blake2s(agent_id.encode("utf-8")).digest()
    return tuple(int.from_bytes(digest[8 * i : 8 * i + 8], "little") for i in range(4))  # type: ignore[return-value]


def id_embedding_words(agent_id: str, embed_words: int = 1) -> np.ndarray:
    """First ``8*embed_words`` digest bytes of blake2s(agent_id) as little-endian uint64 words.

    The device divides each byte by 255.0f in float32, as ``_embed_agent_id`` does.
    """
    return np.array(_digest_words(str(agent_id))[:embed_words], dtype=np.uint64)
