#!/bin/bash
# e2e through the search_func_t entry point with PAGEABLE host text (what krep's mmap hands over), by copy-thread count
python - <<'PY'
import ctypes as C, os, sys, time, subprocess
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from krep_b200 import lib
from krep_b200.abi import Params
n = 8 << 30
L = lib.load(); assert L.krep_b200_init(0) == 0
spec = lib.make_spec(bench.SEED, bench.PLANT_SEED, 1 << 20, b"qzXv9Kpw", 0)
t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
L.krep_b200_corpus_generate(C.byref(spec), t.data_ptr(), 0, n, None)
host = t[:n].cpu().numpy()          # pageable
del t; torch.cuda.empty_cache()
p = Params(b"qzXv9Kpw")
for it in range(3):
    t0 = time.perf_counter()
    cnt, pos = lib.search("avx2", p, None, text_ptr=host.ctypes.data, text_len=n)
    dt = time.perf_counter() - t0
    print(f"threads={os.environ.get('KREP_B200_COPY_THREADS','default')} stage_mb={os.environ.get('KREP_B200_STAGE_MB','32')} pageable 8 GiB: {dt*1e3:.1f} ms -> {n/dt/1e9:.1f} GB/s, matches {cnt}", flush=True)
PY
