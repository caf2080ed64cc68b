"""Yardstick only: batched cuFFT (through torch.fft) on the equatorial ring class of nside 2048
-- 4097 rings of 8192 pixels, Hermitian half spectra of 4097 coefficients -- against the hand-written ring kernels on the
same class.  Prints ms and effective GB/s (algorithmic bytes: half spectrum in + pixels out)."""
import sys
import torch
sys.path.insert(0, '.')


def t_ms(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for nside in (2048, 4096):
    nring, n = 2 * nside + 1, 4 * nside
    H = torch.randn(nring, n // 2 + 1, dtype=torch.complex128, device='cuda')
    x = torch.empty(nring, n, dtype=torch.float64, device='cuda')
    ms_c2r = t_ms(lambda: torch.fft.irfft(H, n=n, dim=1, out=x))
    Ho = torch.empty_like(H)
    ms_r2c = t_ms(lambda: torch.fft.rfft(x, dim=1, out=Ho))
    by = H.numel() * 16 + x.numel() * 8
    print('nside %d: %d rings x %d pixels: cuFFT Z2D %.3f ms (%.0f GB/s), D2Z %.3f ms (%.0f GB/s)'
          % (nside, nring, n, ms_c2r, by / ms_c2r / 1e6, ms_r2c, by / ms_r2c / 1e6))
