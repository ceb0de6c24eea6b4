import sys, ctypes, numpy as np, torch
sys.path.insert(0, '/root/repo')
import asr_ttl_mtl_b200 as b
from asr_ttl_mtl_b200 import _native, audio
lib = _native.load()
def dec(k):
    k = k.astype(np.uint32)
    neg = (k & 0x80000000) == 0
    bits = np.where(neg, ~k, k & 0x7fffffff).astype(np.uint32)
    return bits.view(np.float32)
for nm in (80, 128):
    B, L = 32, 480000
    g = torch.Generator('cuda').manual_seed(1234)
    x = (0.1 * torch.randn(B, L, device='cuda', generator=g)).clamp_(-1, 1)
    T = 3000
    plan = audio._plan(0, nm)
    out = torch.empty(B, nm, T, device='cuda')
    ws = torch.zeros(lib.b200mel_workspace_bytes_tiles(B, T), dtype=torch.uint8, device='cuda')
    st = lib.b200mel_logmel_device(plan, x.data_ptr(), 0, B, L, L, None, 0, out.data_ptr(), ws.data_ptr(), 2, 2, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    w = ws.cpu().numpy().view(np.uint32)
    maxk = w[:B]; mink = w[2*B+1:3*B+1]
    off = lib.b200mel_workspace_bytes(B)//4
    tk = w[off:off+B*24*2].reshape(B*24, 2)
    gmax = dec(maxk); 
    tmax = dec(tk[:,0]); tmin = dec(~tk[:,1])
    floor = np.repeat(gmax, 24) - 8
    print(nm, 'status', st, 'gmax range', gmax.min(), gmax.max(), 'tile min range', tmin.min(), tmin.max(), 'tiles needing clamp', int((~(tmin >= floor)).sum()), 'never stored', int((tk[:,0]==0).sum()))
    print('  out min/max', float(out.min()), float(out.max()))
