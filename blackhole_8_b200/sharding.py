"""How work is dealt to the GPUs of one box (one process per GPU, no collective on the data path).

Every ray is independent, so there is nothing to exchange or reduce:
  * a fly-through shards by whole frames: frame k goes to rank k % world (round-robin keeps the
    ranks in step when the camera moves from cheap to expensive views);
  * a single large frame shards by interleaved row stripes: stripe s (stripe_rows rows) goes to rank
    s % world -- interleaving because the cost per pixel is very uneven (disc hits end after a few
    steps, sky misses run all 2 nstep - 1), contiguous bands would not balance.
Rank r's kernel stores its pixels straight into the gathered buffer on GPU 0 (an IPC / peer mapping),
so the "gather" is the kernel's own stores over NVLink.  These helpers are the single definition of
who owns what; the CUDA kernel's tile-origin arithmetic (bh8_kernel.cuh) implements stripe_rows_of().
"""

TILE_ROWS = 8  # kTileH in csrc/bh8_kernel.cuh: stripe_rows must be a multiple of it


def frames_of(n_frames, rank, world):
    """Indices of the frames rank `rank` renders."""
    return list(range(rank, n_frames, world))


def owner_of_frame(k, world):
    return k % world


def stripe_rows_of(height, stripe_rows, rank, world):
    """[(row_begin, row_end), ...] of the row stripes rank `rank` renders of a frame `height` rows high."""
    if stripe_rows <= 0 or stripe_rows % TILE_ROWS:
        raise ValueError("stripe_rows must be a positive multiple of %d" % TILE_ROWS)
    out = []
    n_stripes = (height + stripe_rows - 1) // stripe_rows
    for s in range(rank, n_stripes, world):
        out.append((s * stripe_rows, min(height, (s + 1) * stripe_rows)))
    return out


def ring_slot_offset(step, rank, world, slots, frame_bytes):
    """Byte offset of rank's frame for `step` inside GPU 0's frame ring (slots x world frames)."""
    return ((step % slots) * world + rank) * frame_bytes


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores next to GPU `device_index` (NVML's ideal CPU affinity).

    One process per GPU reads its frames back into pinned host memory; with first-touch placement
    the pinned pages land on the NUMA node the process runs on, so a process that runs on the other
    socket sends every frame across the inter-socket link.  Returns the CPU set it bound to, or None
    when NVML is not available or reports nothing usable (the process is then left where it is)."""
    import os
    try:
        import pynvml as nv
        nv.nvmlInit()
        handle = nv.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # no NVML / not permitted: placement stays as the launcher chose it
        return None
