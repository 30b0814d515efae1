"""Throughput of cuFFT's strided batched transforms along the slow axis of a [NY][B] array (the y-transform of the
distributed Poisson stage): real-to-complex against complex-to-complex on pairs of columns (development aid)."""
import torch, time
def t(f, n=10):
    f(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for NY,B in ((512, 64*1024), (256, 64*256), (512, 256*1024)):
    x=torch.randn(NY,B,dtype=torch.float64,device='cuda')
    z=torch.view_as_complex(x.view(NY,B//2,2))
    ms_r=t(lambda: torch.fft.rfft(x,dim=0))
    ms_c=t(lambda: torch.fft.fft(z,dim=0))
    gb_r=(x.numel()*8 + (NY//2+1)*B*16)/1e9
    gb_c=(2*x.numel()*8)/1e9
    y=torch.fft.rfft(x,dim=0)
    ms_ir=t(lambda: torch.fft.irfft(y,n=NY,dim=0))
    ms_ic=t(lambda: torch.fft.ifft(z,dim=0))
    print(NY,B,"rfft %.3f ms %.0f GB/s | c2c %.3f ms %.0f GB/s | irfft %.3f ms | ifft %.3f ms"%(ms_r,gb_r/ms_r*1e3,ms_c,gb_c/ms_c*1e3,ms_ir,ms_ic),flush=True)
