import ctypes, torch, sys
sys.path.insert(0, "/root/repo")
from dynamic_llava_b200._lib import lib
L=lib(); P=lambda t: None if t is None else ctypes.c_void_p(t.data_ptr()); ST=lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def timeit(fn, n_rot, iters=30, warm=3):
    for i in range(warm): fn(i % n_rot)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i % n_rot)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters*1e3
for name,B,T,H,rate in [("7B B=8 keep0.2",8,32,4096,0.2),("7B B=16 keep1.0",16,32,4096,1.0),("13B cfg4 B=4",4,1473,5120,0.2),("13B B=32",32,1473,5120,0.2)]:
    n_img,s0=576,5; Pn=T-1+n_img; k=int(n_img*rate); Pp=Pn-n_img+k
    nb=2*B*Pp*H*2
    n_rot=max(2,int(300e6//(B*Pn*H*2))+1)
    hid=[torch.randn(B,Pn,H,device="cuda",dtype=torch.bfloat16) for _ in range(n_rot)]
    out=[torch.empty(B,Pp,H,device="cuda",dtype=torch.bfloat16) for _ in range(n_rot)]
    logits=torch.randn(B,n_img,2,device="cuda",dtype=torch.bfloat16)
    keep=torch.empty(B,k,dtype=torch.int64,device="cuda"); pos=torch.empty(B,Pp,dtype=torch.int32,device="cuda"); cu=torch.empty(B+1,dtype=torch.int32,device="cuda")
    res={}
    for R in (0,8,16,28,37,56,74,111,148,296):
        if R>Pp: continue
        us=timeit(lambda i: L.select_compact(P(logits),None,P(hid[i]),P(out[i]),P(keep),P(pos),P(cu),B,Pn,H,s0,n_img,k,R,ST()),n_rot)
        res[R]=(round(us,1), round(nb/us/1e3))
    print(name, res, flush=True)
    del hid,out
