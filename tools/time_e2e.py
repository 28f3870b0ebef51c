"""End-to-end time of ssak_ctc_loss_host on a workload for several sub-batch counts: python tools/time_e2e.py 1k 4 8 16"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ssak_b200, bench
name = sys.argv[1]
lib = ssak_b200.lib()
lp, tg, il, tl, cells = bench.make_batch(name, 1236)
T, B, V = lp.shape
lp_pin, grad_pin = lp.pin_memory(), torch.empty_like(lp).pin_memory()
tg32, il32, tl32 = tg.to(torch.int32).contiguous(), il.to(torch.int32), tl.to(torch.int32)
nll = torch.empty(B).pin_memory()
ctx = C.c_void_p(); assert lib.ssak_context_create(0, C.byref(ctx)) == 0
for ns in sys.argv[2:]:
    os.environ["SSAK_HOST_SUBBATCHES"] = ns
    def step():
        rc = lib.ssak_ctc_loss_host(ctx, lp_pin.data_ptr(), T, B, V, tg32.data_ptr(), tg32.shape[1], il32.data_ptr(), tl32.data_ptr(), 0, 1, None, nll.data_ptr(), grad_pin.data_ptr())
        assert rc == 0, rc
    step(); step()
    t0 = time.perf_counter()
    for _ in range(5): step()
    print(name, "sub-batches", ns, "ms", round((time.perf_counter() - t0) / 5 * 1e3, 3), flush=True)
