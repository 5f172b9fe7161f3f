import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from garment_pattern_estimation_b200 import _lib
lib = _lib.load()
dev = torch.device('cuda:0')
torch.manual_seed(0)
rows, m, n = 16, 128, 16
a = torch.randn(rows, m, device=dev); b = torch.randn(rows, n, device=dev)
want = a.double().t() @ b.double()
ws = torch.full((int(lib.nt_gemm_tn_workspace_bytes()) // 4,), 7.0, device=dev)
out = torch.zeros(m, n, device=dev)
p = lambda t: ctypes.c_void_p(t.data_ptr())
rc = lib.nt_gemm_tn(p(a), a.stride(0), m, p(b), b.stride(0), n, rows, None, 0, 0, None, 1, 1, p(out), out.stride(0), p(ws), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print('rc', rc, lib.nt_last_error())
tile = ws[:128 * 16].view(128, 16)
print('ws first tile: sevens', int((tile == 7).sum()), 'zeros', int((tile == 0).sum()), 'absmax', float(tile.abs().max()))
print(tile[:3, :6]); print(want[:3, :6])
print('out absmax', float(out.abs().max()))
