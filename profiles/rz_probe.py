import torch, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from mujoco_reinforcement_learning_b200 import _lib
lib = _lib.load()
DEV='cuda'
def tc(A,B,M,N,K,split=1):
    C = torch.empty(M,N,device=DEV)
    _lib.check(lib.b200ppo_debug_tc_gemm(_lib.ptr(A),_lib.ptr(B),_lib.ptr(C),M,N,K,0,0,128,split,_lib.stream_ptr()),"x")
    return C
def sp(A,B,M,N,K,split=1):
    C = torch.empty(M,N,device=DEV)
    _lib.check(lib.b200ppo_debug_gemm_split(_lib.ptr(A),_lib.ptr(B),_lib.ptr(C),None,M,N,K,0,0,split,_lib.stream_ptr()),"x")
    return C
g = torch.Generator().manual_seed(0)
for K in (64, 256, 1024, 4096, 16384):
    M=N=256
    A = torch.randn(M,K,generator=g).bfloat16().float(); B = torch.randn(N,K,generator=g).bfloat16().float()
    ref = A.double()@B.double().T
    C = tc(A.to(DEV),B.to(DEV),M,N,K).cpu().double()
    C32 = (A.to(DEV)@B.to(DEV).T).cpu().double()
    e = (C-ref); rel = e.abs().max()/ref.abs().max()
    big = ref.abs() > ref.abs().max()*0.3
    bias = (e*ref.sign())[big].mean()/ref.abs()[big].mean()
    e32=(C32-ref).abs().max()/ref.abs().max()
    print(f"bf16-exact K={K:6d}: tc err {rel:.2e} signed bias (toward +|C|) {bias:+.2e} | torch fp32 matmul err {e32:.2e}")
    # positive data: coherent sums
    A2=A.abs(); B2=B.abs(); ref2=A2.double()@B2.double().T
    C2=tc(A2.to(DEV),B2.to(DEV),M,N,K).cpu().double()
    print(f"      positive data: rel err mean {((C2-ref2)/ref2).mean():+.2e} max {((C2-ref2)/ref2).abs().max():.2e}")
for K in (64, 376, 1024, 4096):
    M=N=256
    A = torch.randn(M,K,generator=g); B = torch.randn(N,K,generator=g)
    ref = A.double()@B.double().T
    C = sp(A.to(DEV),B.to(DEV),M,N,K).cpu().double()
    e=(C-ref); rel=e.abs().max()/ref.abs().max()
    big = ref.abs() > ref.abs().max()*0.3
    bias = (e*ref.sign())[big].mean()/ref.abs()[big].mean()
    print(f"split K={K:6d}: err {rel:.2e} signed bias {bias:+.2e}")
