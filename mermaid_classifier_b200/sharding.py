"""Multi-GPU partitioning of the extraction / scoring paths: independent units, no collective.

Images go round-robin over ranks -- ``rank = image_index % world`` -- the rule the reference uses
to spread source ids over processing jobs (``/root/reference/scripts/launch_processing.py:59-66``).
Scoring rows are split into contiguous blocks.  One process per GPU; weights are replicated.
"""

from __future__ import annotations

from typing import Any, Sequence


def images_for_rank(n_images: int, rank: int, world: int) -> list[int]:
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_images, world))


def rows_for_rank(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row block ``[lo, hi)`` of ``rank`` (sizes differ by at most one)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return (n_rows * rank) // world, (n_rows * (rank + 1)) // world


def merge_rank_outputs(per_rank: Sequence[dict[int, Any]], n_images: int) -> list[Any]:
    """Reassemble per-image results gathered from the ranks into image order."""
    merged: dict[int, Any] = {}
    for d in per_rank:
        for k, v in d.items():
            if k in merged:
                raise ValueError(f"image {k} was produced by two ranks")
            merged[k] = v
    missing = [i for i in range(n_images) if i not in merged]
    if missing:
        raise ValueError(f"images {missing[:8]}... were produced by no rank")
    return [merged[i] for i in range(n_images)]
