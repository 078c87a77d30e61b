#!/usr/bin/env python
"""bench.py — benchmark of recgraph_b200 (contract in the task statement, tier framing 4).

Headline workload (BASELINE.json configs[1], "C2"): `-m 2` affine-gap adaptive-banded POA, synthetic 100 kbp graph,
10 000 reads of 1 kbp at 5 % error, CLI-default scoring (M=2 X=4 O=4 E=2 b=1 f=0.01). One step = one pass of the hot
path (DP + traceback kernels) over one batch. BASELINE.json's metric is "reads/s and GCUPS per -m mode", so the line also
carries `modes`: C3 (`-m 5`, 32 paths, 10 kbp graph, 10 000 x 2 kbp reads) and C4 (`-m 9`, 64 paths, 5 kbp graph, 1 kbp
2-breakpoint mosaic reads, R=4 r=0.1 B=1), each with its own roofline, CPU baseline and end-to-end figure.

Multi-GPU: ONE read set of N x reads-per-GPU reads (one seed) is cut into contiguous cost-balanced shards
(recgraph_b200.shard.partition), the graph is replicated per device, every rank aligns its shard with no collective on
the data path, and the GAF text is gathered on rank 0 in input order and compared with the text one GPU produces for the
whole set. Per-GPU work is fixed as N grows: "scaling": "weak", as the task statement prescribes for a sharded path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c5] [--reads R] ...
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODE = 2
METRIC = "reads/s, -m 2 (affine-gap banded POA), synthetic 100 kbp graph, 10k x 1 kbp reads"


def build_workload(args, world):
    """The global read set: world x args.reads reads from ONE seed (rank r aligns its contiguous shard of it)."""
    from recgraph_b200 import synth
    g = synth.make_graph(args.graph_bp, 8, seed=1)
    reads = synth.make_reads(g, args.reads * world, args.read_len, err=0.05, seed=3)
    return g, reads


def clocks_sampler(stop, out, device_index):
    """SM clock + throttle reasons during the timed region (the profiling recipe's clocks line). NVML in-process when
    nvidia_ml_py is importable: forking nvidia-smi from a process that holds a >100 GB CUDA address space stalls the
    launching thread for ~10 ms per sample; the nvidia-smi query is the fallback."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = device_index
        if vis and all(t.strip().isdigit() for t in vis.split(",")):
            idx = int(vis.split(",")[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        bits = [(pynvml.nvmlClocksEventReasonHwSlowdown, 3), (pynvml.nvmlClocksEventReasonHwThermalSlowdown, 4),
                (pynvml.nvmlClocksEventReasonSwThermalSlowdown, 5), (pynvml.nvmlClocksEventReasonSwPowerCap, 6)]
        while not stop.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                f = [str(sm), str(mx), "", "", "", "", ""]
                for bit, pos in bits:
                    f[pos] = "Active" if (rs & bit) else "Not Active"
                out.append(f)
            except Exception:
                pass
            stop.wait(0.2)
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(device_index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 7:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(float(s[0])) for s in samples)
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in samples:
        for k, nm in enumerate(names):
            if s[3 + k].lower().startswith("active"):
                reasons.add(nm)
    return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(samples[0][1])), "reasons": sorted(reasons),
            "samples": len(samples)}


def cpu_oracle_run(gfa_path, reads, procs, per_proc, tmpdir, tag, mode=MODE, extra=(), env=None):
    """Time the CPU oracle (oracle/_build/recgraph_oracle, a faithful single-threaded port of the reference, which cannot
    be compiled here) with one process per host core over read shards. Returns (reads, seconds)."""
    from tests import oracle_lib
    oracle_lib.build()
    exe = os.path.join(ROOT, "oracle", "_build", "recgraph_oracle")
    from recgraph_b200 import synth
    files = []
    n = 0
    for c in range(procs):
        shard = reads[c * per_proc:(c + 1) * per_proc]
        if not shard:
            break
        fa = os.path.join(tmpdir, f"{tag}_{c}.fa")
        with open(fa, "w") as f:
            f.write(synth.fasta(shard))
        n += len(shard)
        files.append(fa)
    t0 = time.perf_counter()
    ps = [subprocess.Popen([exe, "-m", str(mode)] + list(extra) + [fa, gfa_path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                           env=env) for fa in files]
    for p in ps:
        p.wait()
    dt = time.perf_counter() - t0
    if any(p.returncode != 0 for p in ps):
        raise RuntimeError("oracle process failed")
    return n, dt


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores. The Rust reference
    cannot be built in this image (no cargo/rustc, un-vendored crates), so this times the oracle port."""
    if rank != 0:
        return
    g, reads = build_workload(args, 1)
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as d:
        gfa = os.path.join(d, "g.gfa")
        open(gfa, "w").write(g.gfa())
        per_core = 1
        env = dict(os.environ, RGO_PRED32="1")   # > 65 535 rows: outside the reference's 16-bit predecessor domain
        for w in range(args.warmup):
            cpu_oracle_run(gfa, reads[:cores], cores, per_core, d, f"w{w}", env=env)
            if args.warmup > 1:
                break  # one warm-up pass is enough to page the binary in; keep the run bounded
        tot_n, tot_t = 0, 0.0
        for k in range(args.steps):
            lo = (k * cores * per_core) % max(1, len(reads) - cores * per_core)
            n, dt = cpu_oracle_run(gfa, reads[lo:lo + cores * per_core], cores, per_core, d, f"s{k}", env=env)
            tot_n += n
            tot_t += dt
    value = tot_n / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} reads per step (1 per core, one oracle process per core, graph load included), {args.steps} steps"},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def path_graph_counts(g):
    """Rows, path-rows (sum over rows of the paths through them) and group-rows (sum over rows of incoming path edges) of a
    synthetic graph: the units of SURVEY 8d's pathwise work figures."""
    seg_len = [len(s) for s in g.segs]
    path_rows = sum(seg_len[s - 1] for p in g.paths for s in p)
    edges = set()
    for p in g.paths:
        prev = 0
        for s in p:
            edges.add((prev, s))
            prev = s
    indeg = {}
    for a, b in edges:
        indeg[b] = indeg.get(b, 0) + 1
    used = set(s for p in g.paths for s in p)
    group_rows = sum(indeg[s] + (seg_len[s - 1] - 1) for s in used)
    return sum(seg_len) + 2, path_rows, group_rows


def traffic_of(kernel):
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get(kernel), t.get("_source")
    except Exception:
        return None, None


def mode_entries(device, int_peak_gops, hbm_peak, steps, with_cpu):
    """C3 (-m 5) and C4 (-m 9) to the same contract as the headline: device-timed value with inputs resident, end-to-end
    through rg_align_batch with host buffers, INT32 + HBM rooflines from SURVEY 8d's per-unit work, CPU baseline."""
    import numpy as np
    from recgraph_b200 import Aligner, synth
    out = {}
    for key, mode, bp, paths, nreads, rlen, err, mosaic, sc, extra, kern in [
            ("m5_c3", 5, 10000, 32, 10000, 2000, 0.05, 0, {}, [], "k_pathwise_tr_m5"),
            ("m9_c4", 9, 5000, 64, 10000, 1000, 0.02, 2, dict(base_rec_cost=4, multi_rec_cost=0.1, rec_band_width=1.0),
             ["-R", "4", "-r", "0.1", "-B", "1"], "k_pathwise_tr_m9")]:
        g = synth.make_graph(bp, paths, seed=1)
        reads = synth.make_reads(g, nreads, rlen, err=err, seed=3, mosaic_breaks=mosaic)
        al = Aligner(device)
        al.load_gfa_text(g.gfa())
        al.set_scoring(**sc)
        n_rows, path_rows, group_rows = path_graph_counts(g)
        codes, off = al.pack_reads(reads)
        al.upload(codes, off)
        for _ in range(3):
            al.align_staged(mode)
        kms = 0.0
        launches = 0
        t0 = time.perf_counter()
        for _ in range(steps):
            al.align_staged(mode)
            ms, nl, _c = al.kernel_stats()
            kms += ms
            launches += nl
        dt = time.perf_counter() - t0
        res = al.fetch()
        bad = sum(1 for i in range(res.n_reads) if res.reads[i].status & ~(1 | 16))
        al.align_packed(mode, codes, off)
        e0 = time.perf_counter()
        for _ in range(steps):
            r2 = al.align_packed(mode, codes, off)
        de = time.perf_counter() - e0
        dirs = 2 if mode >= 8 else 1
        cols = sum(len(r) for r in reads)   # L - 1 per read
        path_cells = dirs * path_rows * cols
        group_cells = dirs * group_rows * cols
        ops = 2.0 * path_cells + 5.0 * group_cells
        ksec = kms * 1e-3 / steps
        traffic, tsrc = traffic_of(kern)
        entry = {
            "workload": f"-m {mode}, synthetic {bp} bp graph, {paths} paths, {nreads} reads x {rlen} bp, {int(err * 100)}% error"
                        + (", reads from 2-breakpoint path mosaics, R=4 r=0.1 B=1" if mosaic else ""),
            "value": nreads * steps / dt, "unit": "reads/s", "steps": steps, "ms_per_step": 1e3 * dt / steps,
            "kernel_ms_per_step": kms / steps, "gpu_launches": int(launches), "bad_status": bad,
            "gcups_rows_x_columns": dirs * (n_rows - 2) * cols / ksec / 1e9,
            "path_cells_per_s": path_cells / ksec,
            "e2e": {"value": nreads * steps / de, "unit": "reads/s", "h2d_bytes_per_step": int(codes.nbytes + off.nbytes),
                    "d2h_bytes_per_step": int(nreads * 80 + int(r2.n_runs_total) * 8), "ms_per_step": 1e3 * de / steps},
            "roofline": {"bound": "int32_alu", "achieved": ops / ksec / 1e9, "peak": int_peak_gops, "unit": "Gop/s",
                         "frac": ops / ksec / 1e9 / int_peak_gops,
                         "ops_per_unit": "2 per path-cell + 5 per group-cell (SURVEY 8d)", "path_cells_per_step": path_cells,
                         "group_cells_per_step": group_cells, "kernel": "k_pathwise_tr (csrc/pathwise_tr.cu)",
                         "note": "the kernel transports path scores through single-edge rows instead of recomputing them, so it "
                                 "executes far fewer than 2 instructions per path-cell: algorithmic ops / time is what is reported"},
            "roofline_hbm": {"bound": "hbm", "achieved": 0.25 * path_cells / ksec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": 0.25 * path_cells / ksec / 1e9 / hbm_peak, "bytes_per_unit": "0.25 B per path-cell (SURVEY 8d)",
                             "traffic": traffic, "traffic_source": tsrc},
        }
        al.close()
        if with_cpu:
            cores = os.cpu_count() or 1
            procs = max(1, min(cores, 8))   # the oracle holds the reference's n x L x P tensors: ~3 GB per process
            with tempfile.TemporaryDirectory() as d:
                gfa = os.path.join(d, "g.gfa")
                open(gfa, "w").write(g.gfa())
                n, t = cpu_oracle_run(gfa, reads[:procs], procs, 1, d, key, mode=mode, extra=extra)
            entry["cpu_baseline"] = {"value": n / t, "unit": "reads/s", "cores": procs, "kind": "port",
                                     "sample": f"first {n} reads of the workload, one single-threaded oracle process each "
                                               f"({t:.1f} s wall, graph load included; {cores} host cores)"}
        out[key] = entry
    return out


def workload_config(args, world):
    return {"workload": f"C2: -m 2, synthetic {args.graph_bp} bp graph (SNP/indel bubbles), {args.reads} reads x "
                        f"{args.read_len} bp per GPU, 5% error, M=2 X=4 O=4 E=2 b=1 f=0.01",
            "reads_per_gpu": args.reads, "global_reads": args.reads * world, "graph_bp": args.graph_bp, "read_len": args.read_len,
            "l2_policy": "inputs larger than L2: every step rewrites > 100 GB of traceback work-space, nothing of a "
                         "previous step can stay in the 126 MB L2",
            "parallelism": "one global read set cut into contiguous cost-balanced shards, graph replicated per GPU, no "
                           "data-path collective, GAF gathered on rank 0 in input order"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c2")
    ap.add_argument("--reads", type=int, default=10000)
    ap.add_argument("--read-len", type=int, default=1000)
    ap.add_argument("--graph-bp", type=int, default=100000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-modes", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.config == "c5":
        from tools import bench_c5
        bench_c5.main(args, rank, local_rank, world)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (recgraph_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from recgraph_b200 import Aligner, shard
    t_gen0 = time.perf_counter()
    g, all_reads = build_workload(args, world)
    gfa_text = g.gfa()
    lo, hi = shard.partition([float(len(r)) for r in all_reads], world)[rank]
    reads = all_reads[lo:hi]
    al = Aligner(local_rank)
    t_load0 = time.perf_counter()
    al.load_gfa_text(gfa_text)
    graph_load_ms = 1e3 * (time.perf_counter() - t_load0)
    al.set_scoring()
    codes, off = al.pack_reads(reads)
    # pinned host staging for the end-to-end leg
    codes_pin = torch.from_numpy(codes).pin_memory()
    off_pin = torch.from_numpy(off.view(np.int64)).pin_memory()
    codes_h, off_h = codes_pin.numpy(), off_pin.numpy().view(np.uint64)

    al.upload(codes_h, off_h)
    for _ in range(max(args.warmup, 3)):
        al.align_staged(MODE)
    res = al.fetch()
    cells_per_step = sum(res.reads[i].cells for i in range(res.n_reads))
    bad = sum(1 for i in range(res.n_reads) if res.reads[i].status & ~1)
    if bad:
        raise SystemExit(f"{bad} reads ended with an error status")

    # ---- timed: inputs resident in HBM
    samples, stop = [], threading.Event()
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local_rank), daemon=True)
    th.start()
    barrier()
    t0 = time.perf_counter()
    kernel_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        al.align_staged(MODE)
        ms, nl, _c = al.kernel_stats()
        kernel_ms += ms
        launches += nl
    barrier()
    t1 = time.perf_counter()
    stop.set()
    th.join()
    dt = t1 - t0

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H of records and runs)
    al.align_packed(MODE, codes_h, off_h)  # warm
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        r2 = al.align_packed(MODE, codes_h, off_h)
    barrier()
    e1 = time.perf_counter()
    de = e1 - e0
    h2d = int(codes_h.nbytes + off_h.nbytes + 4 * len(reads))
    d2h = int(len(reads) * 80 + int(r2.n_runs_total) * 8)

    # ---- GAF text of the shard, gathered on rank 0 in input order (the host-side serial terms, timed)
    tf0 = time.perf_counter()
    text = al.format_gaf_all(MODE, r2, off_h, first_index=lo)
    format_ms = 1e3 * (time.perf_counter() - tf0)
    tg0 = time.perf_counter()
    parts = shard.gather_in_order([text], world, rank)
    gather_ms = 1e3 * (time.perf_counter() - tg0)

    tdev = torch.tensor([dt, de, kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tdev, op=dist.ReduceOp.MAX)
    dt, de, kernel_ms = [float(x) for x in tdev.tolist()]
    total_reads = len(all_reads) * args.steps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    line = None
    if rank == 0:
        gaf_all = "".join(parts)
        multi = {"gaf_lines": gaf_all.count("\n"), "format_ms_rank0": format_ms, "gather_ms": gather_ms,
                 "graph_load_ms_rank0": graph_load_ms, "shard_rank0": [lo, hi]}
        if world > 1:
            # the whole read set on ONE GPU must give the same text as the gathered shards
            c_all, o_all = al.pack_reads(all_reads)
            r_all = al.align_packed(MODE, c_all, o_all)
            multi["equals_single_gpu_text"] = al.format_gaf_all(MODE, r_all, o_all) == gaf_all
        ip = al.int_peak()
        int_peak_gops = max(ip["iadd3_gops"], ip["vimnmx_gops"], 2 * ip["viaddmnmx_gops"])
        cells_s = cells_per_step * args.steps / (kernel_ms * 1e-3)  # per GPU, device time of the dominant kernel
        traffic, tsrc = traffic_of("k_gap_global_blk")
        line = {
            "metric": METRIC, "value": total_reads / dt, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args, world),
            "gcups": cells_per_step * world * args.steps / dt / 1e9,
            "cells_per_step_per_gpu": cells_per_step,
            "e2e": {"value": total_reads / de, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * de / args.steps,
                    "path": "rg_align_batch: host read codes in, numeric records + run lists out (per GPU shard)"},
            "gpu_launches": int(launches),
            "clocks": summarize_clocks(samples),
            "roofline": {"bound": "int32_alu", "achieved": cells_s * 9 / 1e9, "peak": int_peak_gops, "unit": "Gop/s",
                         "frac": cells_s * 9 / 1e9 / int_peak_gops, "ops_per_cell": 9,
                         "peak_source": "rg_int_peak microbenchmark on this GPU: max(IADD3, VIMNMX, 2 x VIADDMNMX) issue rates",
                         "int_peak": ip, "kernel_ms_per_step": kernel_ms / args.steps,
                         "kernel": "k_gap_global_blk<32,2,true> (csrc/poa_gap_blk.cu)", "traffic": traffic, "traffic_source": tsrc,
                         "note": "the binding bound (SURVEY 8d): integer DP, 9 INT32 ops per in-band cell"},
            "roofline_hbm": {"bound": "hbm", "achieved": cells_s * 0.5 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": cells_s * 0.5 / 1e9 / hbm_peak, "peak_source": peak_src, "traffic": traffic,
                             "note": "secondary bound: the kernel writes 4 trace bit-planes = 0.5 B per cell (SURVEY 8d budgets 1 B)"},
            "multi_gpu": multi,
        }
        if world == 1:
            # end to end through the CLI entry point: FASTA + GFA text in, GAF text out (parse + flatten + upload + kernels +
            # traceback + text), the same call the `recgraph` binary makes
            try:
                from recgraph_b200 import run_cli, synth
                with tempfile.TemporaryDirectory() as d:
                    gfa, fa = os.path.join(d, "g.gfa"), os.path.join(d, "r.fa")
                    open(gfa, "w").write(gfa_text)
                    open(fa, "w").write(synth.fasta(reads))
                    al.close()  # release the headline work-space (~120 GB) first
                    run_cli(["-m", "2", fa, gfa])
                    c0 = time.perf_counter()
                    rc, out, err = run_cli(["-m", "2", fa, gfa])
                    c1 = time.perf_counter()
                    line["e2e_cli"] = {"value": len(reads) / (c1 - c0), "unit": "reads/s", "ms": 1e3 * (c1 - c0), "rc": rc,
                                       "gaf_bytes": len(out), "input_bytes": os.path.getsize(fa) + os.path.getsize(gfa),
                                       "path": "rg_cli_main: FASTA + GFA files in, GAF text out (context creation, parse, "
                                               "graph flattening and upload included)"}
            except Exception as ex:
                line["e2e_cli"] = {"error": str(ex)}
        if world == 1 and not args.no_other_modes:
            try:
                al.close()
                line["modes"] = mode_entries(local_rank, int_peak_gops, hbm_peak, args.steps, not args.no_cpu_baseline)
            except Exception as ex:  # never lose the headline line
                line["modes"] = {"error": str(ex)}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as d:
                gfa = os.path.join(d, "g.gfa")
                open(gfa, "w").write(gfa_text)
                per_core = 2
                n, t = cpu_oracle_run(gfa, reads[:cores * per_core], cores, per_core, d, "cb", env=dict(os.environ, RGO_PRED32="1"))
            line["cpu_baseline"] = {"value": n / t, "unit": "reads/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n} reads of the workload, {per_core} per core, one "
                                              f"single-threaded oracle process per core ({t:.1f} s wall, graph load included)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
