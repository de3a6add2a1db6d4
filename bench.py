#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: GB/s scanned over an HBM-resident synthetic corpus.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gib G] [--workload NAME]

N=1 workload = BASELINE configs[1]: 8-byte literal over 10 GiB of HBM-resident synthetic ASCII, count + all offsets.
For N>1 (launched by torchrun, one rank per GPU) each rank holds its own 10 GiB shard (+halo) of an N x 10 GiB
corpus (weak scaling), scans it, and one NCCL gather brings counts and offsets to rank 0.

A step = one pass of the hot path over the resident corpus: filter+verify kernel, device sort of the occurrence
list, read-back, policy replay into krep's match_result_t (+ the gather for N>1).  Inputs are 10 GiB >> 126 MB L2,
so nothing survives in L2 between steps.  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GIB = 1 << 30
WORKLOADS = {
    # name: (needle, params kwargs, corpus flags, plant period, description)
    "literal8": dict(needle=b"qzXv9Kpw", opts={}, flags=0, period=1 << 20,
                     desc="8-byte literal, case-sensitive, count + all offsets (BASELINE configs[1])"),
    "icase4": dict(needle=b"QzXv", opts=dict(case_sensitive=False), flags=1, period=1 << 20,
                   desc="-i 4-byte literal (BASELINE configs[2])"),
    "word16": dict(needle=b"needleneedle0016", opts=dict(whole_word=True), flags=2, period=1 << 26,
                   desc="-w 16-byte literal, low hit rate (BASELINE configs[4])"),
    "multi1000": dict(needle=b"kqzvxjwpy", opts={}, flags=0, period=1 << 22, multi=1000,
                      desc="1000 patterns of 6-12 bytes (-f), Aho-Corasick result set (BASELINE configs[3])"),
    # side workloads (not BASELINE configs): other regimes of the multi-pattern filter
    "multi1000_8to14": dict(needle=b"kqzvxjwpy", opts={}, flags=0, period=1 << 22, multi=1000, lens=(8, 14),
                            desc="1000 patterns of 8-14 bytes (-f): shortest pattern >= 7, full-word hash filter"),
    "multi1000_i": dict(needle=b"kqzvxjwpy", opts=dict(case_sensitive=False), flags=1, period=1 << 22, multi=1000,
                        desc="1000 patterns of 6-12 bytes, -i"),
}
SEED, PLANT_SEED = 0x5EED0001, 0x5EED0002


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def multi_patterns(n, needle, lens=(6, 12)):
    import random
    rng = random.Random(0x5EED0003)
    alpha = "abcdefghijklmnopqrstuvwxyz"
    pats = {needle}
    while len(pats) < n:
        pats.add("".join(rng.choice(alpha) for _ in range(rng.randint(*lens))).encode())
    return [needle] + sorted(pats - {needle})


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md).  The sampler runs from before the warm-up
    (nvidia-smi takes ~100 ms to produce its first line) and only samples whose timestamp falls inside the
    timed region are reported; if the region was too short to contain one, the nearest samples are used."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, timeout=3.0):
        """nvidia-smi needs ~100 ms before its first line: do not let a short timed region start (and end) before it."""
        t_end = time.time() + timeout
        while self.proc and time.time() < t_end:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                return
            time.sleep(0.01)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[4:8]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        if not rows:
            return out
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= self.t1]
        used = inside
        if not used:  # region shorter than the sampling period: take the samples closest to it
            mid = 0.5 * ((self.t0 or rows[-1][0]) + (self.t1 or rows[-1][0]))
            used = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
        reasons = set()
        for r in used:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(r[1] for r in used), "sm_max_mhz": max(r[2] for r in used),
                "power_w_max": max(r[3] for r in used), "reasons": sorted(reasons), "samples": len(used),
                "samples_inside_timed_region": len(inside)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the stock krep CLI built from /root/reference (oracle/_ref/krep), all host threads
# ------------------------------------------------------------------------------------------------
def write_sample(spec_factory, nbytes, path):
    """Writes corpus bytes [0, nbytes) to `path` using the GPU generator when available, else the host twin."""
    from krep_b200 import lib
    try:
        import torch
        if torch.cuda.is_available():
            L = lib.load()
            t = torch.empty(nbytes + 64, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            L.krep_b200_corpus_generate(C.byref(spec_factory()), t.data_ptr(), 0, nbytes, None)
            lib.check(L)
            t[:nbytes].cpu().numpy().tofile(path)
            del t
            torch.cuda.empty_cache()
            return
    except Exception as e:  # noqa: BLE001
        print(f"[bench] GPU corpus generation unavailable ({e}); using the host twin", file=sys.stderr)
    with open(path, "wb") as f:
        step = 64 << 20
        for off in range(0, nbytes, step):
            f.write(lib.corpus_host(spec_factory(), off, min(step, nbytes - off)))


def krep_cli_cmd(cli, wl, sample_path, pat_file):
    cmd = [cli, "-c", "-o"]                      # -co: count matches (scan + count, no output formatting)
    if not wl["opts"].get("case_sensitive", True):
        cmd.append("-i")
    if wl["opts"].get("whole_word"):
        cmd.append("-w")
    if wl.get("multi"):
        cmd += ["-f", pat_file, sample_path]
    else:
        cmd += [wl["needle"].decode(), sample_path]
    return cmd


def run_cpu_reference(wl_name, wl, sample_bytes, steps, warmup, keep_file=False):
    """Times the unmodified reference on a bounded sample of the workload. -> dict(value GB/s, cores, kind, sample, count, ms)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_oracle
    from krep_b200 import lib
    _, cli = build_oracle.build_ref()
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    sample_path = os.path.join(shm, f"krep_b200_sample_{os.getpid()}.txt")
    pat_file = sample_path + ".pats"
    pats = multi_patterns(wl["multi"], wl["needle"], wl.get("lens", (6, 12))) if wl.get("multi") else None

    def spec_factory():
        return lib.make_spec(SEED, PLANT_SEED, wl["period"], wl["needle"], wl["flags"])

    write_sample(spec_factory, sample_bytes, sample_path)
    if pats:
        with open(pat_file, "wb") as f:
            f.write(b"\n".join(pats) + b"\n")
    cores = os.cpu_count() or 1
    times, count, single = [], None, None
    try:
        if cli:
            cmd = krep_cli_cmd(cli, wl, sample_path, pat_file)
            for it in range(warmup + steps):
                t0 = time.perf_counter()
                r = subprocess.run(cmd, capture_output=True, text=True)
                dt = time.perf_counter() - t0
                if it == 0 and dt * (warmup + steps - 1) > 150.0 and sample_bytes > (64 << 20):
                    # keep the whole reference arm within a few minutes whatever the host: shrink the slice
                    shrink = max(64 << 20, int(sample_bytes * 150.0 / (dt * (warmup + steps - 1))) & ~0xFFFFF)
                    with open(sample_path, "r+b") as f:
                        f.truncate(shrink)
                    sample_bytes = shrink
                    continue  # this run timed the larger slice: not recorded
                if it >= warmup:
                    times.append(dt)
                last = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "0"
                count = int(last.rsplit(":", 1)[-1])
            kind = "reference"
            how = f"stock krep CLI (oracle/_ref/krep, -msse4.2 -mavx2 build) `{' '.join(cmd[1:-1])} FILE`, default threads"
            # SURVEY §8d also asks for the -t 1 figure: one run on the first 256 MiB of the same file (head -c via a slice file)
            try:
                small = min(sample_bytes, 256 << 20)
                small_path = sample_path + ".t1"
                with open(sample_path, "rb") as fi, open(small_path, "wb") as fo:
                    fo.write(fi.read(small))
                cmd1 = cmd[:1] + ["-t", "1"] + cmd[1:-1] + [small_path]
                subprocess.run(cmd1, capture_output=True)
                t0 = time.perf_counter()
                subprocess.run(cmd1, capture_output=True)
                single = {"value": small / (time.perf_counter() - t0) / 1e9, "unit": "GB/s", "sample": f"{small >> 20} MiB, -t 1, one run after one warm-up"}
                os.unlink(small_path)
            except Exception:  # noqa: BLE001
                single = None
        else:
            # compiled reference absent: time the scalar oracle port on one core
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle_util as ou
            from krep_b200.abi import Params
            data = open(sample_path, "rb").read()
            func = "aho_corasick" if pats else "sse42"
            for it in range(warmup + steps):
                p = Params(pats if pats else wl["needle"], count=True, only_matching=True, **wl["opts"])
                t0 = time.perf_counter()
                count, _ = ou.port().run(func, p, data, with_result=False)
                dt = time.perf_counter() - t0
                if it >= warmup:
                    times.append(dt)
            kind, cores = "port", 1
            how = "oracle/krep_oracle.c scalar port, 1 thread"
    finally:
        if not keep_file:
            for pth in (sample_path, pat_file):
                if os.path.exists(pth):
                    os.unlink(pth)
    if not times:  # every recorded slot was consumed by the shrink step (steps == 1, warmup == 0)
        t0 = time.perf_counter()
        subprocess.run(cmd, capture_output=True, text=True)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    extra = {}
    if cli and single:
        extra["single_thread"] = single
    try:
        with open("/proc/cpuinfo") as f:
            extra["cpu_model"] = next((ln.split(":", 1)[1].strip() for ln in f if ln.startswith("model name")), None)
    except OSError:
        pass
    return dict(extra, value=sample_bytes / mean / 1e9, best=sample_bytes / min(times) / 1e9, unit="GB/s", cores=cores, kind=kind,
                sample=f"{sample_bytes / GIB:.2f} GiB slice [0, n) of the same corpus in {shm}; whole-process wall, "
                       f"mean of {len(times)} runs after {warmup} warm-up; {how}",
                count=count, ms=mean * 1e3)


# ------------------------------------------------------------------------------------------------
def main():
    # The contract is ONE JSON line on stdout: keep the real stdout aside and point fd 1 at stderr, so that nothing a
    # native library prints (e.g. NCCL's version banner) can land in front of it.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(out_stream):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gib", type=float, default=10.0, help="corpus GiB per GPU")
    ap.add_argument("--workload", default="literal8", choices=list(WORKLOADS))
    ap.add_argument("--cpu-sample-gib", type=float, default=2.0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    metric = "GB/s scanned (HBM-resident corpus)"
    config = {"workload": f"{wl['desc']}; {args.gib:g} GiB synthetic ASCII per GPU, seed {SEED:#x}, "
                          f"1 planted needle per {wl['period'] >> 10} KiB",
              "needle": wl["needle"].decode(), "bytes_per_gpu": int(args.gib * GIB), "l2": "inputs >> L2 (no flush needed)"}

    if args.impl == "reference":
        if rank != 0:
            return
        sample = int(min(args.cpu_sample_gib, args.gib) * GIB)
        r = run_cpu_reference(args.workload, wl, sample, max(args.steps, 1), args.warmup)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": r["value"], "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread", "cpu_model") if k in r},
            "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "matches_in_sample": r["count"],
        }), file=out_stream)
        return

    import torch
    import torch.distributed as dist
    from krep_b200 import lib, sharding
    from krep_b200.abi import ALGO_AC, ALGO_AVX2, DeviceResult, Params, Shard

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = lib.load()
    assert L.krep_b200_init(local_rank) == 0, L.krep_b200_last_error_string()

    n = int(args.gib * GIB)
    n -= n % 16
    pats = multi_patterns(wl["multi"], wl["needle"], wl.get("lens", (6, 12))) if wl.get("multi") else None
    maxlen = max(map(len, pats)) if pats else len(wl["needle"])
    halo = maxlen + 1
    last = rank == world - 1
    g0, own_len, avail = sharding.shard_bounds(world * n, world, rank, halo)   # weak scaling: n owned bytes per GPU
    assert own_len == n
    spec = lib.make_spec(SEED, PLANT_SEED, wl["period"], wl["needle"], wl["flags"])
    text = torch.empty(avail + 64, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    assert L.krep_b200_corpus_generate(C.byref(spec), text.data_ptr(), g0, avail, sptr) == 0
    prev_byte = -1
    if rank > 0:
        prev_byte = lib.corpus_host(spec, g0 - 1, 1)[0]
    next_byte = -1 if last else lib.corpus_host(spec, g0 + avail, 1)[0]
    torch.cuda.synchronize()

    # -co semantics would not need offsets; the workload asks for count + all offsets -> default mode (track positions)
    params = Params(pats if pats else wl["needle"], **wl["opts"])
    algo = ALGO_AC if pats else ALGO_AVX2      # what select_search_algorithm picks (krep.c:1771): AVX2 entry -> SSE4.2 kernel
    plan = L.krep_b200_plan_create(params.ref(), algo)
    lib.check(L)
    shard = Shard(text.data_ptr(), avail, 0, n, g0, prev_byte, next_byte)
    dev = DeviceResult()
    res = L.krep_b200_match_result_init(1 << 16)

    gatherer = sharding.KeyGatherer(world, rank, "cuda") if world > 1 else None
    checked = [True]   # warm-up steps size the exchange buffers on all ranks; the timed steps then run unchecked

    def step():
        rc = L.krep_b200_scan_shard(plan, C.byref(shard), 1, sptr, C.byref(dev))
        assert rc == 0, L.krep_b200_last_error_string()
        kms = L.krep_b200_last_kernel_ms()
        if world == 1:
            res.contents.count = 0
            total = L.krep_b200_collect(plan, params.ref(), C.byref(dev), res)
            return total, kms
        # N>1: ONE collective — every rank all_gathers [count, sorted keys]; rank 0 reads the rows back and replays the
        # concatenation (rank order = global order), krep_b200/sharding.py
        while True:
            n_mine = int(dev.stored)
            if n_mine <= gatherer.cap:
                L.krep_b200_export_keys(C.byref(dev), gatherer.key_buffer().data_ptr(), n_mine, sptr)
            keys, counts, retry = gatherer.exchange(n_mine, check=checked[0])
            if not retry:
                break
        total = 0
        if rank == 0:
            res.contents.count = 0
            arr = C.cast(keys.data_ptr(), C.POINTER(C.c_uint64))
            total = L.krep_b200_replay(algo, params.ref(), False, arr, keys.numel(), None, world * n, res)
        return total, kms

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3) if args.steps else 0):
        step()
    L.krep_b200_reset_launch_count()
    checked[0] = False
    if rank == 0:
        sampler.wait_first_sample()
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record(stream)
    total = 0
    for _ in range(args.steps):
        total, kms = step()
        kernel_ms.append(kms)
    e1.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = e0.elapsed_time(e1)
    launches = int(L.krep_b200_launch_count())
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        km = torch.tensor([sum(kernel_ms) / len(kernel_ms)], dtype=torch.float64, device="cuda")
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
        kernel_avg_ms = float(km.item())
    else:
        kernel_avg_ms = sum(kernel_ms) / len(kernel_ms)
    ms_per_step = elapsed_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9
    match_count = int(total)
    first = [(res.contents.positions[i].start_offset, res.contents.positions[i].end_offset)
             for i in range(min(3, res.contents.count))] if rank == 0 else []

    # ---- e2e: same metric through the search_func_t entry point with HOST buffers (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e:
        host = torch.empty(avail, dtype=torch.uint8, pin_memory=True)
        host.copy_(text[:avail])
        torch.cuda.synchronize()
        entry = "aho_corasick" if pats else "avx2"
        fn = getattr(L, lib.SEARCH_ENTRIES[entry])
        if pats:
            params.struct.ac_trie = L.krep_b200_ac_trie_build(params.ref())
        L.krep_b200_set_only_matching(False)

        def e2e_step():
            res.contents.count = 0
            c = fn(params.ref(), C.c_void_p(host.data_ptr()), avail, res)
            lib.check(L)
            # ownership by start offset: matches that start in the halo belong to the next rank
            own = sum(1 for i in range(res.contents.count) if res.contents.positions[i].start_offset < n) \
                if not last else int(c)
            return own

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        own = 0
        for _ in range(args.e2e_steps):
            own = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            c = torch.tensor([own], dtype=torch.int64, device="cuda")
            dist.all_reduce(c)
            own = int(c.item())
        e2e = {"value": world * n / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": world * avail,
               "d2h_bytes_per_step": 8 * world + 8 * own, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "api": lib.SEARCH_ENTRIES[entry] + "(params, pinned host text, len, match_result_t*)",
               "matches": own, "agrees_with_device_path": own == match_count}
        params.struct.ac_trie = None
        del host

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    achieved = n / (kernel_avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            tr = json.load(open(tpath)).get(args.workload)
            if tr:
                traffic = tr["dram_bytes_per_launch"] * (n / tr["corpus_bytes"])
        except Exception:  # noqa: BLE001
            pass
    out = {
        "metric": metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": dict(config, filter=L.krep_b200_plan_filter_name(plan).decode(),
                                            parallelism=f"{world} shard(s), owned by match start, halo {halo} B"),
        "matches": match_count, "first_matches": first,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel_ms": kernel_avg_ms,
                     "algorithmic_bytes_per_launch": n,
                     "note": "achieved = corpus bytes of one shard / mean scan-kernel duration (CUDA events on the launching stream, inside the timed region; max over ranks)"},
        "gpu_launches": launches, "clocks": clocks,
    }
    if e2e:
        out["e2e"] = e2e
    if not args.no_cpu and world == 1:
        try:
            r = run_cpu_reference(args.workload, wl, int(min(args.cpu_sample_gib, args.gib) * GIB), 3, 1)
            out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread", "cpu_model") if k in r}
            out["cpu_baseline"]["matches_in_sample"] = r["count"]
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(out), file=out_stream)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
