import os, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); local=int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
n=59_000_000
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/iters
x=torch.randn(n,device=dev)
t=timeit(lambda: dist.all_reduce(x))
if rank==0: print("nccl all_reduce 236MB ms", t, flush=True)
try:
    import torch.distributed._symmetric_memory as sm
    gname=dist.group.WORLD.group_name
    sm.enable_symm_mem_for_group(gname)
    y=sm.empty(n, dtype=torch.float32, device=dev); y.normal_()
    hdl=sm.rendezvous(y, group=gname)
    ops=[o for o in dir(torch.ops.symm_mem)]
    if rank==0: print("symm ops", [o for o in ops if not o.startswith('_')], "multicast", getattr(hdl,'has_multicast_support',None), flush=True)
    for name in ("multimem_all_reduce_","two_shot_all_reduce_","one_shot_all_reduce"):
        try:
            op=getattr(torch.ops.symm_mem,name)
            t=timeit(lambda: op(y,"sum",gname))
            if rank==0: print(name,"ms",t,flush=True)
        except Exception as e:
            if rank==0: print(name,"failed",repr(e)[:200],flush=True)
except Exception as e:
    if rank==0: print("symm_mem failed", repr(e)[:300], flush=True)
dist.destroy_process_group()
