"""bench.py --config c5 — BASELINE.json configs[4], the read-sharded sweep: synthetic 1 Mbp graph with 128 haplotype paths,
reads of 1-10 kbp, modes 2 / 5 / 9, on 1 / 2 / 4 / 8 GPUs (reads sharded, graph replicated, no collective).

The literal workload (10^6 reads) is ~10^18 path-cell updates for mode 5 alone and outside what the reference can
run at all (F3: 16-bit predecessors need n <= 65 535 rows for mode 2; F6: the n x L x P score tensor of modes 5 / 9 would be
5 TB per read), so a STATED SAMPLE per mode is timed and reported as reads/s. What bounds each sample:
  mode 2  reads of 1 000-1 023 bases only: the register-blocked kernel holds a read's columns in one warp (<= 1 024
          columns); longer reads go through the generic striped kernel, which is ~10x slower per cell and is not timed here.
          A global alignment against 1 Mbp costs ~n x L / 2 = 5e8 cells for a 1 kbp read (the band stays full once the
          read is consumed, utils.rs:55-66).
  mode 5  the whole 1-10 kbp range (256 / 384-thread column classes of the score-transport kernel); reads in flight are
          bounded by the 2-bit leader-move trace (n_groups x L / 4 bytes per read: 3.3 GB at 10 kbp).
  mode 9  NOT timed on the 1 Mbp graph: best_alignment's pair reduction visits, for every surviving column, every surviving
          forward node against all n reverse nodes (the reference's loop is O(n^2 L), pathwise_alignment_recombination.rs:
          808-864); the pruning that makes it cheap at C4 (5 kbp) does not bound it at n = 10^6, and the per-(row, column)
          maxima are n x L x 16 bytes per read (16 GB at 1 kbp). The sweep point reported for mode 9 is a 20 kbp, 128-path
          graph with 1 kbp mosaic reads (the largest of this repo's timed mode-9 graphs).
"""
import json
import os
import time


def main(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from recgraph_b200 import Aligner, synth, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (recgraph_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    g = synth.make_graph(args.graph_bp if args.graph_bp != 100000 else 1000000, 128, seed=1)
    gfa_text = g.gfa()
    t_graph = time.perf_counter() - t0
    al = Aligner(local_rank)
    t0 = time.perf_counter()
    al.load_gfa_text(gfa_text)
    t_load = time.perf_counter() - t0
    n_rows, n_segs, P = al.graph_info()
    rng = np.random.default_rng(3)

    def reads_of(lengths, seed, mosaic=0, err=0.05):
        out = []
        for k, ln in enumerate(lengths):
            out += synth.make_reads(g, 1, int(ln), err=err, seed=seed + k, mosaic_breaks=mosaic)
        return out

    per_gpu = {"m2": 592, "m5": 48, "m9": 296}
    samples = {
        "m2": (2, reads_of(rng.integers(1000, 1024, size=per_gpu["m2"] * world), 1000), {},
               "reads of 1 000-1 023 bases (register-blocked kernel: one warp holds the read's columns)"),
        "m5": (5, reads_of(rng.integers(1000, 10001, size=per_gpu["m5"] * world), 2000), {},
               "reads of 1-10 kbp, uniform"),
    }
    # mode 9 on its own, smaller graph (see the module docstring)
    g9 = synth.make_graph(20000, 128, seed=1)
    al9 = Aligner(local_rank)
    al9.load_gfa_text(g9.gfa())
    m9_reads = synth.make_reads(g9, per_gpu["m9"] * world, 1000, err=0.02, seed=3, mosaic_breaks=2)
    out_modes = {}
    samples["m9"] = (9, m9_reads, dict(base_rec_cost=4, multi_rec_cost=0.1, rec_band_width=1.0),
                     "1 kbp reads from 2-breakpoint path mosaics, R=4 r=0.1 B=1, on a 20 kbp / 128-path graph (NOT the 1 Mbp graph)")
    for key, (mode, reads, sc, what) in samples.items():
        lo, hi = shard.partition([float(len(r)) for r in reads], world)[rank]
        mine = reads[lo:hi]
        if key == "m9":
            al.close()
            al = al9
        al.set_scoring(**sc)
        codes, off = al.pack_reads(mine)
        al.upload(codes, off)
        entry = {"mode": mode, "sample": f"{len(reads)} reads ({len(reads) // world} per GPU): {what}"}
        try:
            al.align_staged(mode)   # warm-up: work-space allocation
            barrier()
            t0 = time.perf_counter()
            kms = 0.0
            for _ in range(args.steps):
                al.align_staged(mode)
                kms += al.kernel_stats()[0]
            barrier()
            dt = time.perf_counter() - t0
            res = al.fetch()
            bad = sum(1 for i in range(res.n_reads) if res.reads[i].status & ~(1 | 16))
            cells = sum(res.reads[i].cells for i in range(res.n_reads))
            tdev = torch.tensor([dt, kms, float(cells), float(bad)], dtype=torch.float64, device="cuda")
            if world > 1:
                tmax = tdev.clone()
                dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tdev, op=dist.ReduceOp.SUM)
                dt, kms = float(tmax[0]), float(tmax[1])
                cells, bad = float(tdev[2]), float(tdev[3])
            entry.update({"value": len(reads) * args.steps / dt, "unit": "reads/s", "ms_per_step": 1e3 * dt / args.steps,
                          "kernel_ms_per_step": kms / args.steps, "gcups": cells * args.steps / dt / 1e9, "bad_status": int(bad),
                          "bases_per_step": sum(len(r) for r in reads)})
        except Exception as ex:   # a mode that does not fit must not lose the others
            entry["error"] = str(ex)
        out_modes[key] = entry
    if rank == 0:
        line = {
            "metric": "reads/s per -m mode on the C5 sweep sample (1 Mbp graph, 128 paths, 1-10 kbp reads)", "unit": "reads/s",
            "value": out_modes["m5"].get("value"), "n_gpus": world, "steps": args.steps, "warmup": 1, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "C5: 1 Mbp synthetic pangenome graph, 128 haplotype paths, reads of 1-10 kbp, modes 2 / 5 / 9; "
                                   "a stated sample per mode and GPU (the literal 10^6 reads are ~10^18 path-cell updates)",
                       "rows": n_rows, "segments": n_segs, "paths": P, "graph_gfa_mb": len(gfa_text) / 1e6,
                       "graph_generate_s": t_graph, "graph_parse_flatten_upload_s": t_load},
            "modes": out_modes,
            "reference_cannot_run": {
                "mode 2": "F3: the traceback predecessor is 16 bits (bitfield_path.rs:39-44): graphs above 65 535 rows corrupt or "
                          "panic; largest graph it can run: 60 kbp (tests/test_gpu_fullsize.py compares that size literally)",
                "modes 5 / 9": "F6: the full n x L x P i32 tensor (pathwise_alignment_semiglobal.rs:17) is 1e6 x 1e4 x 128 x 4 B = 5 TB "
                               "per read, twice for mode 9 plus an n x n displacement matrix (4 TB) and an O(n^2 L) pair loop"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
