"""north_star: "tensor cores are used only if a DFT-as-GEMM stage for the short subsampled lengths beats the butterfly
path".  The candidate: the 126 phi leaves of a signal (63 packed pairs + S0 = 64 complex rows): periodised spectrum of
128 bins -> 128-point inverse transform -> the 75 samples [27, 102) that are kept, i.e. ONE real GEMM
    [rows x 256] . [256 x 150 (padded to 160)]      in 3xTF32 (fp32-class accuracy, like the phase path's stage B).

Measured here, on the device, for 16 384 signals (1 048 576 rows):
  * the GEMM as three TF32 cuBLAS products (a hand-written tcgen05 kernel cannot be faster than the library at this
    plain shape; our own tcgen05 kernel sustains 315 TFLOP/s on stage B) -- time, TFLOP/s, accuracy against float64;
  * the butterfly path for the same transforms: the interpreter's 128-point inverse passes on the same rows, as tile
    jobs from global memory (LOADC -> radix 8, radix 16 -> STOREC): an UPPER bound of their cost inside the fused
    cascade, where the rows never leave shared memory.
Prints one JSON line.  Not on the product path."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'vae-teb_b200')):
    sys.path.insert(0, p)

import numpy as np      # noqa: E402
import torch            # noqa: E402


def main():
    from tebscat import _lib
    from tebscat.large import LargeDevicePlan
    from tebscat.schedule import bitrev_indices
    dev = torch.device('cuda')
    n_sig, rows_per_sig, L, keep0, keep1 = 16384, 64, 128, 27, 102
    rows = n_sig * rows_per_sig
    g = torch.Generator(device='cuda').manual_seed(3)
    z = torch.randn(rows, L, 2, device=dev, generator=g)                    # spectra (bit-reversed bin order), re/im
    # inverse DFT matrix restricted to the kept samples, acting on bit-reversed bins
    br = bitrev_indices(L)
    k = br[:, None].astype(np.float64)
    n = np.arange(keep0, keep1)[None, :].astype(np.float64)
    W = np.exp(2j * np.pi * k * n / L)                                      # [bin slot, sample]
    # real form: [re, im] row vector times [[Wr, Wi], [-Wi, Wr]] -> [re, im] of the samples
    Bm = np.zeros((2 * L, 2 * (keep1 - keep0) + 10), np.float64)            # padded to 160 columns
    nk = keep1 - keep0
    Bm[0::2, 0:nk], Bm[0::2, nk:2 * nk] = W.real, W.imag
    Bm[1::2, 0:nk], Bm[1::2, nk:2 * nk] = -W.imag, W.real
    A = z.reshape(rows, 2 * L)
    Bt = torch.from_numpy(Bm).to(dev)

    def split(t):                                                            # TF32 head (round to nearest) and tail
        hi = (t.view(torch.int32) + 0x1000 & ~0x1fff).view(torch.float32)
        return hi, t - hi
    B32 = Bt.float()
    Bh, Bl = split(B32)
    torch.backends.cuda.matmul.allow_tf32 = True

    def gemm3():
        Ah, Al = split(A)
        return Ah @ Bh + (Al @ Bh + Ah @ Bl)

    for _ in range(2):
        out = gemm3()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = gemm3()
    e1.record()
    torch.cuda.synchronize()
    gemm_ms = e0.elapsed_time(e1) / 5
    # products only (the split and the two adds are elementwise passes a fused kernel would not pay)
    Ah, Al = split(A)
    e0.record()
    for _ in range(5):
        p1, p2, p3 = Ah @ Bh, Al @ Bh, Ah @ Bl
    e1.record()
    torch.cuda.synchronize()
    mm_ms = e0.elapsed_time(e1) / 5
    sub = slice(0, 4096)
    ref = A[sub].double() @ Bt
    err = float((out[sub].double() - ref).norm() / ref.norm())
    flops = 3 * 2.0 * rows * (2 * L) * Bm.shape[1]

    # butterfly path: the interpreter's 128-point inverse transform on the same rows (tile jobs from global memory)
    class _P:
        tile_lengths = [7]
        arena = np.zeros(4, np.float32)
    ctx = LargeDevicePlan(_P(), 0)
    buf = z.clone().contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib = _lib.load()
    for _ in range(2):
        _lib.check(lib.tebscat_large_fft(ctx.handle, ctypes.c_void_p(buf.data_ptr()), rows, 7, 1, st))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        _lib.check(lib.tebscat_large_fft(ctx.handle, ctypes.c_void_p(buf.data_ptr()), rows, 7, 1, st))
    e1.record()
    torch.cuda.synchronize()
    bfly_ms = e0.elapsed_time(e1) / 5
    print(json.dumps({
        'tool': 'dft_gemm_experiment', 'signals': n_sig, 'rows': rows, 'gemm_shape': [rows, 2 * L, Bm.shape[1]],
        'gemm_3xtf32_ms': gemm_ms, 'gemm_products_only_ms': mm_ms, 'gemm_tflops_executed': flops / (mm_ms * 1e-3) / 1e12,
        'gemm_rel_l2_vs_float64': err, 'gemm_flops_per_signal': flops / n_sig,
        'butterfly_flops_per_signal': rows_per_sig * 5 * L * 7, 'butterfly_tile_jobs_from_global_ms': bfly_ms,
        'butterfly_bytes_moved_gb': 2 * rows * L * 8 / 1e9,
        'note': 'butterfly figure includes a global-memory round trip of every row that the fused cascade does not make'}))


if __name__ == '__main__':
    main()
