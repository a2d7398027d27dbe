"""numpy restatement of the counter-based RNG stream behind ``torch.randn`` on CUDA
(oracle; test infrastructure).

The reference draws its noise with ``torch.randn(shape, device)`` /
``torch.randn_like`` (sr3 diffusion.py:174,194; indi.py:67,82).  On a CUDA device
those resolve to Philox4x32-10 (cuRAND ``curand_init(seed, thread, offset)`` +
``curand_normal4``) inside ATen's grid-stride kernel
(torch/include/ATen/native/cuda/DistributionTemplates.h:34-90, 444-453):

    block = 256, grid = min(SMs * (maxThreadsPerSM // 256), ceil(numel / 256))
    thread idx draws float4 #k from counter (offset/4 + k, subsequence idx)
    lane ii of that float4 lands at element  idx + grid*256*ii + k*grid*256*4
    generator offset advances by ((numel-1) // (256*grid*4) + 1) * 4

Integer part (Philox) is bit-exact here; the Box-Muller transform uses fp32 numpy
``log``/``sin``/``cos`` and therefore agrees with the device (``logf`` +
``__sincosf``) only to a few ulp - tests compare with a 2e-6 absolute tolerance,
while device-vs-``torch.randn`` is checked bit-exactly on the GPU box.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr (n,4) uint32, key (2,) uint32 -> (n,4) uint32 (curand_philox4x32_x.h:160-192)."""
    c = ctr.astype(np.uint32).copy()
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    for r in range(10):
        p0 = M0 * c[:, 0].astype(np.uint64)
        p1 = M1 * c[:, 2].astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
        c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
        if r != 9:
            with np.errstate(over="ignore"):
                k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
                k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c


def box_muller(x: np.ndarray, y: np.ndarray):
    """curand_normal.h:70-87 (device branch), fp32."""
    inv = np.float32(2.3283064e-10)
    inv2pi = np.float32(2.3283064e-10) * np.float32(6.2831855)
    # uint32 -> fp32 conversion rounds first; the multiply-add is contracted to one FMA on device
    xf, yf = x.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64)
    u = (xf * np.float64(inv) + np.float64(inv / np.float32(2))).astype(np.float32)
    v = (yf * np.float64(inv2pi) + np.float64(inv2pi / np.float32(2))).astype(np.float32)
    s = np.sqrt(np.float32(-2.0) * np.log(u)).astype(np.float32)
    return (np.sin(v) * s).astype(np.float32), (np.cos(v) * s).astype(np.float32)


def launch_grid(numel: int, sm_count: int, max_threads_per_sm: int = 2048) -> int:
    return min(sm_count * (max_threads_per_sm // 256), (numel + 255) // 256)


def offset_increment(numel: int, sm_count: int, max_threads_per_sm: int = 2048) -> int:
    grid = launch_grid(numel, sm_count, max_threads_per_sm)
    return ((numel - 1) // (256 * grid * 4) + 1) * 4


def randn_like_cuda(numel: int, seed: int, offset: int, sm_count: int,
                    max_threads_per_sm: int = 2048) -> np.ndarray:
    """The flat fp32 tensor ``torch.randn(numel, device='cuda')`` yields for generator
    state (seed, offset) on a device with ``sm_count`` SMs."""
    assert offset % 4 == 0
    grid = launch_grid(numel, sm_count, max_threads_per_sm)
    nthreads = grid * 256
    rounds = (numel - 1) // (nthreads * 4) + 1
    out = np.empty(rounds * nthreads * 4, dtype=np.float32)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    idx = np.arange(nthreads, dtype=np.uint64)
    for k in range(rounds):
        cnt = offset // 4 + k
        ctr = np.stack([np.full(nthreads, cnt & 0xFFFFFFFF, dtype=np.uint32),
                        np.full(nthreads, (cnt >> 32) & 0xFFFFFFFF, dtype=np.uint32),
                        (idx & _MASK).astype(np.uint32),
                        (idx >> np.uint64(32)).astype(np.uint32)], axis=1)
        r = philox4x32_10(ctr, key)
        n0, n1 = box_muller(r[:, 0], r[:, 1])
        n2, n3 = box_muller(r[:, 2], r[:, 3])
        base = k * nthreads * 4
        for ii, n in enumerate((n0, n1, n2, n3)):
            out[base + ii * nthreads: base + (ii + 1) * nthreads] = n
    return out[:numel]
