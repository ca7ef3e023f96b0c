"""Why does the host path (e2e) not gain from the single batch?  Launch counts and times of both paths, free memory."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import corpus
b2 = importlib.import_module("zip-ada_b200")
n = 1 << 30
dev = torch.device("cuda", 0)
d_in = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
d_in[:n] = corpus.workload("markov", n, 0x5EED0001, torch, dev)
cap = int(b2.lib().b2_bound(n)) + 1024 * (n // 40000 + 16)
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.copy_(d_in[:n])
h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize(); torch.cuda.empty_cache()
def free(): return round(torch.cuda.mem_get_info()[0] / 2**30, 1)
def run(enc, host, tag):
    enc.reset_stats(); torch.cuda.synchronize(); t0 = time.perf_counter()
    if host: ln = enc.encode_ptr(h_in.data_ptr(), n, n, h_out.data_ptr(), cap)
    else: ln = enc.encode_device_ptr(d_in.data_ptr(), n, n, d_out.data_ptr(), cap)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = enc.stats()
    print(tag, "wall_ms", round(dt * 1e3, 1), "call_ms", round(st.call_ms, 1), "launches", int(st.kernel_launches), "scatter_launches", int(st.scatter_launches), "sort_ms", round(st.sort_ms, 1), "free_GiB", free(), flush=True)
print("free at start", free())
with b2.Encoder(9, 0) as enc:
    enc.set_timing(1)
    for i in range(2): run(enc, False, "dev%d" % i)
    for i in range(3): run(enc, True, "host%d" % i)
    run(enc, False, "dev_again")
with b2.Encoder(9, 0) as enc:
    enc.set_timing(1)
    for i in range(3): run(enc, True, "fresh_host%d" % i)
