"""Round-robin partition of a batch of independent frames over the GPUs of one box.

BASELINE.json north_star: "a batch of independent frames is partitioned round-robin across
the 8 GPUs of one box (no NCCL - frames share nothing)".  The reference's only parallel
driver splits ROWS of one frame over threads (libswscale/swscale.c:1645-1679); across
devices the natural unit is the frame.  No data-path collective exists; ranks only meet
at the timing barrier of bench.py.
"""


def frames_for_rank(n_frames, rank, world):
    """Global frame indices handled by `rank`: i with i % world == rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def frames_per_rank(n_frames, world):
    """How many frames every rank gets (the first n_frames % world ranks get one more)."""
    base, extra = divmod(n_frames, world)
    return [base + (1 if r < extra else 0) for r in range(world)]
