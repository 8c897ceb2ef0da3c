#!/usr/bin/env python
"""bench.py -- primers/s of the `dicey hunt` hot path on the synthetic 3 Gb reference.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/_ref)

Workload (BASELINE.json metric: "primers/sec (3 Gb ref, edit-dist 1)"): 24 x 125 Mb iid-uniform
ACGT records, 1,000,000 20-mers per GPU (half planted with <= 1 edit, half random),
edit distance 1, both strands, dicey defaults (-m 1000 -x 10000).  One step = one pass of the whole
hot path (prepare, neighbour search, antichain filter, locate, NW verify) over the batch.
The index is replicated per GPU and the primer batch is sharded by rank (weak scaling: every rank
runs 1 M primers); `value` counts the primers of all ranks.

Keys beyond the base contract:
  roofline      k_search (the dominant kernel): algorithmic bytes of SURVEY.md 8(d),
                32 R + 32 L + 4 H + X per primer taken from the instrumented reference run on the
                same index, x primers per launch / CUDA-event duration of that kernel
  cpu_baseline  the reference (oracle/_ref/dicey_ref: verbatim SDSL / neighbors.h / needle.h) on the
                host cores, on a bounded sample of the same primers over the same 3 Gb index
  e2e           the same metric through the C ABI (dg_hunt_batch) with pinned host buffers:
                H2D of the queries and D2H of every hit record inside the timed region
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "dicey_ref")      # the reference's own code ("kind": "reference")
PORT_BIN = os.path.join(ROOT, "oracle", "dicey_oracle")          # the plain C++ restatement ("kind": "port")


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_binary():
    """(path, kind) of the CPU comparator: the reference compiled from its own sources when it was
    built (oracle/_ref), otherwise the oracle port; (None, None) if neither exists."""
    if os.path.exists(REF_BIN):
        return REF_BIN, "reference"
    if os.path.exists(PORT_BIN):
        return PORT_BIN, "port"
    return None, None
SEED = 42


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--primers", type=int, default=1_000_000, help="primers per GPU per step")
    ap.add_argument("--length", type=int, default=20)
    ap.add_argument("--distance", type=int, default=1)
    ap.add_argument("--hamming", action="store_true")
    ap.add_argument("--nrec", type=int, default=24)
    ap.add_argument("--reclen", type=int, default=125_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ref-sample", type=int, default=0, help="reference arm: primers per step (0 = auto)")
    ap.add_argument("--fm9-dir", default=None)
    ap.add_argument("--ramp", type=int, default=60, help="extra untimed resident steps after the W warm-up steps (clock ramp); "
                    "a fixed count, identical on every rank, because every step holds a collective")
    ap.add_argument("--ramp-e2e", type=int, default=25, help="the same for the end-to-end arm")
    ap.add_argument("--skip-resident", action="store_true", help="(experiments) run the end-to-end arm only")
    ap.add_argument("--gather-to-host", action="store_true", help="N > 1: rank 0 also copies the gathered coordinate table "
                    "of all ranks to its host inside every end-to-end step")
    return ap.parse_args()


class Watchdog:
    """A multi-rank run that stops making progress names itself instead of sitting in a collective
    until torch's watchdog fires: every rank records its last milestone; if none is reached for
    `limit` seconds the rank prints where it is and exits."""

    def __init__(self, rank: int, limit: float = 150.0):
        self.rank, self.limit, self.where, self.t = rank, limit, "start", time.time()
        self._stop = threading.Event()
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def mark(self, where: str):
        self.where, self.t = where, time.time()
        if os.environ.get("BENCH_PROGRESS"):
            print(f"[bench rank {self.rank}] {where}", file=sys.stderr, flush=True)

    def _run(self):
        while not self._stop.wait(2.0):
            if time.time() - self.t > self.limit:
                print(f"[bench rank {self.rank}] no progress for {self.limit:.0f} s after milestone '{self.where}': giving up",
                      file=sys.stderr, flush=True)
                os._exit(3)

    def stop(self):
        self._stop.set()


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md), read through NVML
    by the timing thread itself right after each step's synchronisation.  A concurrent poller (an
    nvidia-smi child, or an NVML thread) takes driver locks while kernels are being launched and
    was measured to stretch individual steps by 2-180 ms on this box; two NVML reads between steps
    cost ~0.05 ms.  Falls back to one nvidia-smi query after the run if NVML is unavailable."""

    def __init__(self, device: int):
        self.device, self.rows, self.h = device, [], None
        if os.environ.get("BENCH_CLOCK", "") == "off":
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES if it is a plain list of ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = device
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[device])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            for _ in range(3):   # the first NVML reads are slow: take them before the timed region
                self.sample()
            self.rows = []
        except Exception:
            self.h = None

    def sample(self):
        if self.h is None:
            return
        try:
            self.rows.append((float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)), int(self.reasons_fn(self.h))))
        except Exception:
            pass

    def start(self):
        self.rows = []

    def stop(self) -> dict:
        if self.h is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits", "-i",
                                      str(self.device)], capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": ["nvml unavailable: one nvidia-smi sample after the run"]}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.rows))
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max, "samples": len(sm), "reasons": reasons}


def make_primers(args, rank: int) -> np.ndarray:
    from dicey_b200 import synth
    return synth.primers_fast(SEED, args.nrec, args.reclen, args.primers, args.length, args.distance,
                              not args.hamming, rng_seed=7 + rank)


def write_sample(path: str, primers: np.ndarray):
    with open(path, "wb") as f:
        f.write(b"\n".join(bytes(row) for row in primers) + b"\n")


def run_ref(fm9: str, rec: str, qfile: str, args, threads: int, counters: bool, records: str | None = None) -> dict:
    cmd = [cpu_binary()[0], "hunt", fm9, rec, qfile, "-d", str(args.distance), "--threads", str(threads)]
    if records:
        cmd += ["--records", records]
    if args.hamming:
        cmd.append("-n")
    if counters:
        cmd.append("--counters")
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def workload_name(args) -> str:
    gb = args.nrec * args.reclen / 1e9
    mode = "hamming" if args.hamming else "edit"
    return (f"dicey hunt: {args.primers} {args.length}-mers per GPU, {mode}-distance {args.distance}, both strands, "
            f"{gb:.3g} Gb synthetic reference ({args.nrec} x {args.reclen} bp)")


def build_index(args, device: int):
    from dicey_b200.api import Index
    t0 = time.time()
    ix = Index.build_synthetic(SEED, args.nrec, args.reclen, device)
    return ix, time.time() - t0


def ensure_fm9(ix, args) -> tuple[str, str, float]:
    d = args.fm9_dir or ("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    fm9 = os.path.join(d, f"dicey_b200_bench_{args.nrec}x{args.reclen}.fm9")
    rec = fm9 + ".rec.tsv"
    t0 = time.time()
    ix.write_fm9(fm9)
    with open(rec, "w") as f:
        for i in range(args.nrec):
            f.write(f"chr{i + 1}\t{args.reclen}\n")
    return fm9, rec, time.time() - t0


def cpu_leg(fm9, rec, primers, args, cores, seconds):
    """Reference on the host cores over a bounded sample; returns (primers/s, sample size, counters)."""
    tmp = tempfile.mkdtemp(prefix="dicey_b200_cpu_")
    probe_n = min(len(primers), 64 * cores)
    qf = os.path.join(tmp, "probe.txt")
    write_sample(qf, primers[:probe_n])
    r = run_ref(fm9, rec, qf, args, cores, False)
    rate = max(r["queries_per_s"], 1e-9)
    n = int(min(len(primers), max(probe_n, rate * seconds)))
    write_sample(qf, primers[:n])
    rec_out = os.path.join(tmp, "ref.records.tsv")
    r = run_ref(fm9, rec, qf, args, cores, True, rec_out)
    return r, n, rec_out


def parity_check(ix, params, primers, n, ref_records: str) -> dict:
    """canonical dump of the hit records (push order and sorted order, alignment
    strings included) of the CPU sample, reference vs CUDA path on the same 3 Gb index; SHA-256 of both."""
    import hashlib
    res = ix.hunt(primers[:n], params)
    mine = res.records_tsv(params, primers[:n])
    ref = "".join(l for l in open(ref_records) if not l.startswith("W\t"))
    a, b = hashlib.sha256(mine.encode()).hexdigest(), hashlib.sha256(ref.encode()).hexdigest()
    out = {"primers": int(n), "hits": int(len(res.hits)), "sha256_gpu": a, "sha256_reference": b, "equal": a == b}
    if a != b:
        ml, rl = mine.splitlines(), ref.splitlines()
        out["first_difference"] = next(((i, x, y) for i, (x, y) in enumerate(zip(ml, rl)) if x != y), (min(len(ml), len(rl)), "", ""))
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref_bin, ref_kind = cpu_binary()
    if ref_bin is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/dicey_ref nor oracle/dicey_oracle has been built"}))
        return 0
    cores = os.cpu_count() or 1
    ix, _ = build_index(args, int(os.environ.get("LOCAL_RANK", "0")))
    fm9, rec, _ = ensure_fm9(ix, args)
    ix.close()
    primers = make_primers(args, 0)
    tmp = tempfile.mkdtemp(prefix="dicey_b200_ref_")
    sample = args.ref_sample
    if not sample:
        qf = os.path.join(tmp, "probe.txt")
        write_sample(qf, primers[:32 * cores])
        rate = max(run_ref(fm9, rec, qf, args, cores, False)["queries_per_s"], 1e-9)
        sample = int(max(32 * cores, min(len(primers) // max(1, args.steps + args.warmup), rate * 8.0)))
    loop_s, done = 0.0, 0
    for step in range(args.warmup + args.steps):
        lo = (step * sample) % max(1, len(primers) - sample)
        qf = os.path.join(tmp, f"step{step}.txt")
        write_sample(qf, primers[lo:lo + sample])
        r = run_ref(fm9, rec, qf, args, cores, False)
        if step >= args.warmup:
            loop_s += r["loop_s"]
            done += r["queries"]
    value = done / loop_s if loop_s > 0 else 0.0
    sample_desc = f"{sample} primers per step of the same primer set and 3 Gb index; per-primer loop only (index load excluded)"
    line = {
        "impl": "reference", "metric": "primers/sec (3 Gb ref, edit-dist 1)", "value": value, "unit": "primers/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * loop_s / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args), "threads": cores},
        "cpu_baseline": {"value": value, "unit": "primers/s", "cores": cores, "cpu_model": cpu_model(), "kind": ref_kind, "sample": sample_desc},
        "e2e": {"value": value, "unit": "primers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main_b200(args):
    import torch
    import torch.distributed as dist
    from dicey_b200.api import HuntParams

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (dicey_b200 has no CPU path)")
    torch.cuda.set_device(local)
    # (one rank holds no collective that could hang: a long limit there lets profilers replay kernels in peace)
    dog = Watchdog(rank, float(os.environ.get("BENCH_WATCHDOG_S", "150" if world > 1 else "3600")))
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    dog.mark("process group up")
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    ix, build_s = build_index(args, local)
    dog.mark("index built")
    info = ix.info()
    params = HuntParams(distance=args.distance, hamming=args.hamming)
    primers = make_primers(args, rank)
    dog.mark("primers made")
    nq = primers.shape[0]
    # pinned host copies of the inputs
    pin = torch.empty(primers.size, dtype=torch.uint8, pin_memory=True)
    pin.numpy()[:] = primers.reshape(-1)
    off_pin = torch.empty(nq + 1, dtype=torch.int64, pin_memory=True)
    off_pin.numpy()[:] = np.arange(nq + 1, dtype=np.int64) * args.length
    seqs = (pin.numpy(), off_pin.numpy().view(np.uint64))

    stream = torch.cuda.ExternalStream(ix.stream(), device=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: queries staged once, every kernel per step; with more than
    # one rank each step ends with the path's one exchange step: dg_allgather_hits, ONE ncclAllGather of
    # the 16-byte hit records issued on the index stream behind the step's kernels (C ABI, dg_comm.cu)
    from dicey_b200 import shard
    comm = shard.comm_from_process_group(ix) if world > 1 else None
    dog.mark("communicator up")
    qbase = rank * nq
    if comm is not None:
        comm.set_query_base(qbase)     # peer mode: the producing kernels store global query ids into every rank's table
    gathered = [0]

    def resident_step(batch):
        batch.run()
        if comm is not None:
            _, _, counts = comm.allgather_hits(batch, qbase)   # synchronises the index stream
            gathered[0] = int(counts.sum())
        batch.summary()          # synchronises the index stream; collects the stage timings

    batch = ix.stage(seqs, params)
    # W warm-up steps, then a FIXED number of further untimed steps: the GPU idles while the primers are
    # generated on the host and needs a few hundred ms of load to come back to its boost clock.  (A
    # wall-clock bound here would let ranks run different numbers of steps -- and every step holds a
    # collective.)
    for i in range(2 if args.skip_resident else args.warmup + args.ramp):
        resident_step(batch)
    dog.mark("resident warm-up done")
    ix.profile(True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    profs = []
    for _ in range(1 if args.skip_resident else args.steps):
        resident_step(batch)
        profs.append(ix.last_profile())
        sampler.sample()         # right after the step's synchronisation: the clock it ran at
    e1.record(stream)
    barrier()
    dog.mark("resident arm timed")
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    nhits, ncand = batch.summary()
    ix.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * nq * args.steps / (ms / 1e3)
    batch.free()

    # ---------------- end-to-end arm: host buffers through the C ABI: dg_hunt_batch (H2D of the queries,
    # every kernel, D2H of every hit record and alignment of this rank's shard to this rank's host).
    # With several ranks the 16-byte hit records (coordinates, strand, distance, global query id) are
    # then all-gathered on the device (dg_allgather_hits), so every rank holds the coordinate table of
    # the whole batch in HBM; with --gather-to-host rank 0 also copies that table to its host.
    def e2e_step():
        res = ix.hunt(seqs, params)          # H2D, every kernel, D2H of every record (chunk-pipelined)
        if comm is None:
            return res, 0
        comm.allgather_hits(None, qbase)
        extra = 0
        if args.gather_to_host and rank == 0:
            extra = comm.fetch_table().nbytes
        return res, extra

    for _ in range(min(args.warmup, 2) + args.ramp_e2e):
        res, _ = e2e_step()
    dog.mark("e2e warm-up done")
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    step_ms = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        res, extra = e2e_step()
        step_ms.append(round(1e3 * (time.perf_counter() - ts), 3))
        d2h = res.transfer_bytes + extra      # counted by the library from the copies it issued
        e2e_hits = len(res.records)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    dog.mark("e2e arm timed")
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * nq * args.steps / e2e_s
    # equal-length batches upload the sequence bytes only (the device derives the offsets); the offsets
    # array is read by the host side of the call
    h2d = int(pin.numel())
    # where the end-to-end time goes (untimed extra passes through the split form of the same call)
    phases = []
    for _ in range(3):
        t0 = time.perf_counter(); bt = ix.stage(seqs, params)
        t1 = time.perf_counter(); bt.run()
        t2 = time.perf_counter(); r2 = bt.fetch()
        t3 = time.perf_counter(); bt.free(); del r2
        t4 = time.perf_counter()
        phases.append([round(1e3 * (b - a), 3) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))])

    # ---------------- CPU reference beside it (rank 0, single GPU run only) + roofline numerator
    cpu = None
    work = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu and cpu_binary()[0]:
        dog.limit = 1200.0      # no collectives from here on: the CPU leg may take its time
        dog.mark("cpu leg")
        try:
            fm9, rec, write_s = ensure_fm9(ix, args)
            cores = os.cpu_count() or 1
            r, n, ref_records = cpu_leg(fm9, rec, primers, args, cores, args.cpu_seconds)
            parity = parity_check(ix, params, primers, n, ref_records)
            cpu = {"value": r["queries_per_s"], "unit": "primers/s", "cores": cores, "cpu_model": cpu_model(), "kind": cpu_binary()[1],
                   "sample": f"first {n} primers of the batch on the same 3 Gb index ({r['loop_s']:.1f} s loop, index load excluded)"}
            work = {k: r[k] / r["queries"] for k in ("R", "L", "H", "X", "strings", "steps")} if "R" in r else None
            try:
                os.remove(fm9); os.remove(fm9 + "_check")
            except OSError:
                pass
        except Exception as e:  # the bench line must still be printed
            cpu = {"value": None, "unit": "primers/s", "cores": os.cpu_count(), "kind": cpu_binary()[1], "sample": f"failed: {e}"}
    if work is None:
        # SURVEY.md 8(d) expectation for 20-mers at edit distance 1 on 3 Gb (used only when the
        # reference could not be run, e.g. under torchrun N > 1; the N = 1 run measures it)
        fallback = os.path.join(ROOT, "profiles", "work_counters.json")
        try:
            work = json.load(open(fallback))[f"{'h' if args.hamming else 'e'}{args.distance}"]
        except Exception:
            work = None

    # ---------------- the exchanged table itself (untimed): every rank checks what it holds in HBM after the
    # last end-to-end step -- its own segment against its own result, every segment's query range and order
    exchange_check = None
    if comm is not None:
        ok, gathered_n, err = False, 0, None
        try:
            res_last, _ = e2e_step()
            _, _, cnts = comm.allgather_hits(None, qbase)
            tab = comm.fetch_table()
            off = np.concatenate([[0], np.cumsum(cnts.astype(np.int64))])
            gathered_n = int(off[-1])
            ok = len(tab) == gathered_n
            for r in range(world):
                seg = tab[off[r]:off[r + 1]]
                ok = ok and (len(seg) == 0 or (int(seg["query"].min()) >= r * nq and int(seg["query"].max()) < (r + 1) * nq
                                               and bool((np.diff(seg["query"].astype(np.int64)) >= 0).all())))
            mine = tab[off[rank]:off[rank + 1]]
            recs = res_last.records
            ok = ok and len(mine) == len(recs) and all(bool(np.array_equal(mine[f], recs[f] + (qbase if f == "query" else 0)))
                                                       for f in ("query", "chr", "start", "score", "strand"))
        except Exception as e:   # (every rank still takes part in the reduction below; the bench line is still printed)
            ok, err = False, str(e)
        t_ok = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        exchange_check = {"all_ranks_ok": bool(t_ok.item()), "gathered_records": gathered_n}
        if err:
            exchange_check["error"] = err
    dog.mark("exchange checked")

    ms_search = float(np.mean([p["ms_search"] for p in profs])) if profs else None
    ms_probe = float(np.mean([p["ms_probe"] for p in profs])) if profs else 0.0
    scripts = float(np.mean([p["scripts"] for p in profs])) if profs else 0.0
    roof = None
    if ms_search:
        # The dominant kernel of the step is k_probe_singles (one presence-bitmap probe per neighbour string;
        # its CUDA-event time is taken live: dg_profile.ms_probe).  `achieved` = the DRAM bytes one launch
        # moves (ncu --set full of the same command, profiles/k_search_traffic.json, regenerated by
        # tools/make_traffic_json.py) / that time; `frac` = achieved / the measured copy peak.  Beside it:
        # the algorithmic figure (one 32-byte sector per probe), the probe rate against the random-gather
        # ceiling of tools/gather_bench2.cu (profiles/gather_ceiling.json), DRAM sectors demanded (L1 and L2
        # misses) against sectors fetched, and the SURVEY.md 8(d) number of the reference's own work
        # (32 B x its rank queries) as `reference_work_ratio` -- the tables answer most of those queries
        # without DRAM, so that one says how much work the design avoids, not how close a kernel is to a limit.
        # (k_resolve of one tile runs on a side stream beside the probes of the next tile, and together they are
        # bound by the same DRAM transaction rate: the search stage is reported as one unit -- its CUDA-event time,
        # the DRAM bytes of both kernels)
        split_stage = ms_probe > 0
        dom_ms = ms_search
        dom = "k_probe_singles + k_resolve (search stage, two streams)" if split_stage else "k_search_packed"
        roof = {"bound": "hbm", "kernel": dom, "achieved": None, "peak": peak_gbs, "unit": "GB/s", "frac": None,
                "peak_source": peak_src, "traffic": None, "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / (ms / args.steps),
                "search_stage_ms": ms_search, "probe_kernels_ms": ms_probe, "probes_per_launch": scripts,
                "algorithmic_bytes_per_launch": 32.0 * scripts,
                "algorithmic_achieved": 32.0 * scripts / (dom_ms / 1e3) / 1e9,
                "algorithmic_frac": 32.0 * scripts / (dom_ms / 1e3) / 1e9 / peak_gbs,
                "probes_per_s": scripts / (dom_ms / 1e3)}
        tr = os.path.join(ROOT, "profiles", "k_search_traffic.json")
        if os.path.exists(tr):
            try:
                t_ = json.load(open(tr))
                ks_ = t_.get("kernels", {})
                k_ = None
                if split_stage and "k_probe_singles" in ks_ and "k_resolve" in ks_:
                    a_, b_ = ks_["k_probe_singles"], ks_["k_resolve"]
                    k_ = {"dram_bytes": a_["dram_bytes"] + b_["dram_bytes"], "dram_sectors_read": a_["dram_sectors_read"] + b_["dram_sectors_read"],
                          "l1_sectors": None, "l1_hit_pct": None, "l2_hit_pct": None,
                          "demanded": sum(x["l1_sectors"] * (1 - x["l1_hit_pct"] / 100) * (1 - x["l2_hit_pct"] / 100) for x in (a_, b_)),
                          "ncu_ms": [a_.get("duration_ms"), b_.get("duration_ms")]}
                elif not split_stage:
                    k_ = ks_.get("k_search_packed")
                if (k_ and t_.get("kmer") == info["kmer"] and t_.get("bitmap_k") == info["bitmap_k"] and t_.get("primers") == nq
                        and t_.get("index_device_bytes") == info["device_bytes"] and not args.hamming and args.distance == 1):
                    roof["traffic"] = k_["dram_bytes"]
                    roof["achieved"] = k_["dram_bytes"] / (dom_ms / 1e3) / 1e9
                    roof["frac"] = roof["achieved"] / peak_gbs
                    roof["dram_sectors_fetched"] = k_["dram_sectors_read"]
                    if k_.get("demanded"):
                        roof["dram_sectors_demanded"] = k_["demanded"]
                        roof["ncu_kernel_ms"] = k_["ncu_ms"]
                    elif k_.get("l1_sectors") and k_.get("l1_hit_pct") is not None and k_.get("l2_hit_pct") is not None:
                        roof["dram_sectors_demanded"] = k_["l1_sectors"] * (1 - k_["l1_hit_pct"] / 100) * (1 - k_["l2_hit_pct"] / 100)
                    roof["traffic_source"] = t_.get("source")
            except Exception:
                pass
        if roof["achieved"] is None:   # no capture for this configuration: the algorithmic figure stands in
            roof["achieved"], roof["frac"] = roof["algorithmic_achieved"], roof["algorithmic_frac"]
            roof["note"] = ("no ncu capture for this configuration: achieved = 32 B per probe / kernel time -- an upper bound on the DRAM "
                            "bytes (sibling probes share sectors, L1 / L2 answer many), so frac can exceed 1 and is not a DRAM utilisation")
        gc = os.path.join(ROOT, "profiles", "gather_ceiling.json")
        if os.path.exists(gc):
            try:
                g_ = json.load(open(gc))
                roof["gather_ceiling_per_s"] = g_["gathers_per_s"]
                roof["frac_gather"] = roof["probes_per_s"] / g_["gathers_per_s"]
            except Exception:
                pass
        if work:
            roof["reference_work_ratio"] = 32 * work["R"] * nq / (ms_search / 1e3) / 1e9 / peak_gbs
            roof["reference_bytes_per_primer"] = 32 * work["R"] + 32 * work["L"] + 4 * work["H"] + work["X"]
    if rank == 0:
        stage = {k: float(np.mean([p[k] for p in profs])) for k in ("ms_prepare", "ms_search", "ms_filter", "ms_locate", "ms_verify", "ms_total")}
        line = {
            "metric": "primers/sec (3 Gb ref, edit-dist 1)", "value": value, "unit": "primers/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(args), "parallelism": f"index replicated x{world}, primers sharded by rank" + (", one ncclAllGather of the 16-byte hit records per step on the index stream (dg_allgather_hits)" if world > 1 else ""),
                       "gathered_hits_per_step": gathered[0] if world > 1 else None,
                       "e2e_gather_to_host": bool(args.gather_to_host) if world > 1 else None,
                       "global_primers_per_step": world * nq, "l2": f"inputs larger than L2: random access into {info['device_bytes'] / 1e9:.0f} GB of index tables (the same staged batch every step; a step moves ~8 GB of DRAM traffic through the 126 MB L2)",
                       "kmer_table_K": info["kmer"], "presence_bitmap_K": info["bitmap_k"], "index_device_bytes": info["device_bytes"],
                       "index_build_s": build_s, "hits_per_step": nhits, "candidates_per_step": ncand},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "primers/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                    "records_per_step": int(e2e_hits), "record_bytes": 24,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "phases_ms_stage_run_fetch_free": phases, "hunt_call_ms": step_ms},
            "gpu_launches": int(sum(p["launches"] for p in profs)) + (2 * args.steps if world > 1 else 0),
            "stages_ms": stage, "step_ms_total": [round(p["ms_total"], 3) for p in profs],
            "roofline": roof,
            "exchange_check": exchange_check,
            "cpu_baseline": cpu,
            "parity": parity,
        }
        print(json.dumps(line))
    dog.mark("line printed")
    if comm is not None:
        comm.close()
    ix.close()
    if world > 1:
        dist.destroy_process_group()
    dog.stop()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_b200(a))
