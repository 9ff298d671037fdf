#!/usr/bin/env python
"""bench.py -- molecules/sec at 1000 denoise steps (GEOM-Drugs synthetic), the BASELINE.json metric.

A "step" is ONE body of the reverse-diffusion loop over one synthetic batch: MolDiff.forward (6-block
NodeEdgeNet denoiser) + posterior sampling + (guided workloads) BondPredictor forward + d/dpos backward.
Per-step cost does not depend on t, so
        molecules/sec @ 1000 steps = n_molecules / (1000 * seconds_per_step).

The step is replayed as ONE CUDA graph (MolDiff.graphed_step -- what MolDiff.sample does by default; --no-graph launches kernel by
kernel); tools/full_sample_run.py times a complete 1000-step MolDiff.sample next to this per-step figure.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload guided|unguided|train_fwd] [--batch B] [--simple] [--no-graph]
  python bench.py --impl reference ...      # the CPU oracle port (reference's own op sequence) on host cores

N > 1: launched by torch.distributed.run, one rank per GPU; molecules are sharded (weak scaling, B per
GPU fixed; --strong --batch 2048: ONE batch split by balancing sum n^2), no data-path collective; one NCCL gather of
the final predictions after the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 1000
# as-written reference GEMM FLOPs per directed edge (torch FlopCounterMode)
FLOP_EDGE_DENOISER_FWD = 7.814e6
FLOP_EDGE_BOND_FWD = 8.132e6
FLOP_EDGE_BOND_FWD_BWD = 16.052e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "guided", "unguided", "train_fwd"],
                    help="guided (default, BASELINE config 2) / unguided sampler loop body, or train_fwd = MolDiff.get_loss "
                         "forward + loss (BASELINE config 3: --workload train_fwd --batch 1024)")
    ap.add_argument("--batch", type=int, default=None, help="molecules per GPU (default 256 = BASELINE config 2; "
                                                            "1024 for train_fwd, 32 for --simple)")
    ap.add_argument("--simple", action="store_true",
                    help="BASELINE config 1: sample_MolDiff_simple.yml (train_MolDiff_simple.yml weights), T = 50, B = 32, unguided")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True,
                    help="replay the sampler loop body as one CUDA graph (MolDiff.graphed_step) -- the default, as in MolDiff.sample")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="eager launches (one per kernel)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: --batch is the GLOBAL batch, split over the ranks by balancing sum n^2 (FLOPs)")
    ap.add_argument("--max-size", type=int, default=None,
                    help="every molecule has exactly this many atoms (reference make_data_placeholder(max_size=...); "
                         "BASELINE config 5: --batch 8192 --max-size 29 --workload unguided)")
    ap.add_argument("--cpu-batch", type=int, default=None, help="molecules in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sust=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def ncu_traffic():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of each kernel class, from the committed
    `ncu --set full` capture of the latest round (profiles/rNN_ncu_metrics.json, written by tools/ncu_metrics.py)."""
    for name in ("r02_ncu_metrics.json", "r01_ncu_metrics.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return {k: v.get("traffic_bytes") for k, v in json.load(f).items()}, f"profiles/{name} (ncu --set full, one launch)"
        except Exception:
            continue
    return {}, None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        os.unlink(self.f.name)
        return out


def build_models(workload, device, simple=False):
    from moldiff_b200 import BondPredictor, MolDiff
    from moldiff_b200.config import builtin_config
    torch.manual_seed(0)
    if simple:      # BASELINE config 1: train_MolDiff_simple.yml with diff.num_timesteps overridden to 50 (and nothing else)
        cfg = builtin_config("train/train_MolDiff_simple.yml").model
        cfg["diff"]["num_timesteps"] = 50
        model = MolDiff(cfg, 8, 6).eval()
    else:
        model = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).eval()
    bond = None
    if workload == "guided":
        torch.manual_seed(0)
        bond = BondPredictor(builtin_config("train/train_bondpred.yml").model, 8, 5).eval()
    if device is not None:
        model = model.to(device)
        bond = bond.to(device) if bond is not None else None
    return model, bond


def cpu_step_seconds(workload, B, threads, reps=1):
    """One loop body of the reference algorithm (oracle port: the reference's own unfused op sequence in
    PyTorch CPU) on B molecules; returns seconds per step (min over reps after one warm-up)."""
    from oracle import restatement as R
    torch.set_num_threads(threads)
    model, bond = build_models(workload, None)
    sd = {k: v for k, v in model.state_dict().items()}
    sdb = {k: v for k, v in bond.state_dict().items()} if bond is not None else None
    np.random.seed(2023)
    ph = R.make_data_placeholder(B)
    bn, hei, bh = ph["batch_node"], ph["halfedge_index"], ph["batch_halfedge"]
    N, Eh = len(bn), len(bh)
    g = torch.Generator().manual_seed(1)
    state = (torch.nn.functional.one_hot(torch.randint(0, 8, (N,), generator=g), 8).float(),
             torch.randn(N, 3, generator=g),
             torch.nn.functional.one_hot(torch.randint(0, 6, (Eh,), generator=g), 6).float(),
             R.index_to_log_onehot(torch.randint(0, 8, (N,), generator=g), 8),
             R.index_to_log_onehot(torch.randint(0, 6, (Eh,), generator=g), 6))
    noise = dict(pos=torch.randn(N, 3, generator=g), node=torch.rand(N, 8, generator=g), edge=torch.rand(Eh, 6, generator=g))
    guidance = ("uncertainty", 1e-4) if workload == "guided" else None
    best = float("inf")
    for r in range(reps + 1):
        t0 = time.perf_counter()
        with torch.no_grad():
            R.sample_step(sd, state, bn, hei, bh, 500, B, noise, sd_bond=sdb, guidance=guidance)
        dt = time.perf_counter() - t0
        if r > 0:
            best = min(best, dt)
    return best, N, Eh


def run_reference(args, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.cpu_batch or (32 if workload == "guided" else 64)
    times = []
    _ = cpu_step_seconds(workload, B, threads, reps=0) if args.warmup > 0 else None
    steps = max(1, min(args.steps, 5))
    for _i in range(steps):
        dt, N, Eh = cpu_step_seconds(workload, B, threads, reps=1)
        times.append(dt)
    sec = float(np.mean(times))
    value = B / (T_STEPS * sec)
    line = {
        "impl": "reference", "metric": "molecules/sec at 1000 denoise steps", "value": value, "unit": "molecules/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"sample_MolDiff {workload}, 1 loop body of the 1000-step sampler on a {B}-molecule "
                               f"GEOM-Drugs-sized synthetic batch (N={N}, E={2 * Eh}); CPU oracle port of the reference",
                   "batch": B},
        "cpu_baseline": {"value": value, "unit": "molecules/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} loop bodies at B={B} (N={N}, E={2 * Eh}), per-step cost is linear in E"},
        "e2e": {"value": value, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    workload = args.workload or os.environ.get("MDB_BENCH_WORKLOAD", "guided")
    if args.simple:
        workload = "unguided"
    if args.impl == "reference":
        return run_reference(args, workload if workload != "train_fwd" else "unguided")

    from moldiff_b200 import engine
    from moldiff_b200.placeholder import draw_sizes, make_data_placeholder
    from moldiff_b200.sharding import balanced_shards
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    engine.load_library()
    model, bond = build_models(workload, dev, simple=args.simple)
    guidance = ("uncertainty", 1e-4) if workload == "guided" else None
    T = model.num_timesteps                       # 1000 (50 for --simple): molecules/s = B / (T * seconds per loop body)
    train = workload == "train_fwd"

    B = args.batch or (1024 if train else 32 if args.simple else 256)
    if args.strong and world > 1:                 # ONE batch of B molecules split over the ranks, sum n^2 balanced
        np.random.seed(2023)
        sizes = draw_sizes(B, max_size=args.max_size)
        ph = make_data_placeholder(None, sizes=sizes[balanced_shards(sizes, world)[rank]])
        total_mols, scaling = B, "strong"
        B = int(ph["batch_node"].max()) + 1
    else:                                         # weak scaling: every rank draws its own B molecules
        np.random.seed(2023 + rank)
        ph = make_data_placeholder(B, max_size=args.max_size)
        total_mols, scaling = B * world, "weak"
    host = {k: v.pin_memory() for k, v in ph.items()}
    d = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    N, Eh = len(ph["batch_node"]), len(ph["batch_halfedge"])
    E = 2 * Eh
    torch.manual_seed(2023 + rank)

    graphed = None
    if train:
        # synthetic 'dataset' batch (as tests/golden/make_golden_loss.py): real atom / bond classes, N(0, 1) * 2 positions
        g = torch.Generator().manual_seed(9 + rank)
        host_in = {"node_type": torch.randint(0, 7, (N,), generator=g).pin_memory(),
                   "node_pos": (torch.randn(N, 3, generator=g) * 2.0).pin_memory(),
                   "halfedge_type": torch.randint(0, 5, (Eh,), generator=g).pin_memory()}
        mol = {k: v.to(dev) for k, v in host_in.items()}
        last = {}

        def step(i):
            with torch.no_grad():
                last["loss"] = model.get_loss(mol["node_type"], mol["node_pos"], d["batch_node"], mol["halfedge_type"],
                                              d["halfedge_index"], d["batch_halfedge"], B)
            return last["loss"]
        eager_step = step
    else:
        st = model.sample_begin(B, d["batch_node"], d["halfedge_index"], d["batch_halfedge"])

        def eager_step(i):
            return model.sample_step(st, T - 1 - (i % T), bond_predictor=bond, guidance=guidance)
        step = eager_step
        if args.graph:
            for i in range(2):
                eager_step(i)
            graphed = model.graphed_step(st, bond_predictor=bond, guidance=guidance)

            def step(i):
                return graphed.run(T - 1 - (i % T))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()

    # ---- timed region: device-resident inputs ----
    l2_flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    clocks = ClockSampler(local) if rank == 0 else None
    n0 = engine.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        l2_flush.zero_()                      # flush L2 between timed iterations (working set ~ L2 size)
        ev[i][0].record()
        step(args.warmup + i)
        ev[i][1].record()
    barrier()
    launches = engine.launch_count() - n0
    if graphed is not None:                   # replays do not pass through the library's launch counter
        launches += graphed.launches_per_replay * args.steps
    ms = sum(a.elapsed_time(b) for a, b in ev)
    clk = clocks.stop() if clocks is not None else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank_ms = None
    if dist is not None:
        mine = t_ms.clone() / args.steps
        allt = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine)
        per_rank_ms = [round(float(x.item()), 4) for x in allt]
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    per_pass = 1 if train else T              # train_fwd: molecules per second through one forward + loss pass
    value = total_mols / (per_pass * ms_per_step * 1e-3)

    # ---- e2e: same step through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if train:
        h2d = sum(v.numel() * v.element_size() for v in host_in.values()) + sum(v.numel() * v.element_size() for v in host.values())
        d2h = 16
        res_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        e0.record()
        for i in range(args.steps):
            for k in host_in:
                mol[k] = host_in[k].to(dev, non_blocking=True)
            for k in host:
                d[k] = host[k].to(dev, non_blocking=True)
            out = step(args.warmup + i)
            res_host.copy_(torch.stack([out["loss"], out["loss_pos"], out["loss_node"], out["loss_edge"]]), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e1.record()
    else:
        cur = graphed.st if graphed is not None else st
        keys = ("h_node", "pos", "h_half", "log_node", "log_half")
        host_state = {k: cur[k].detach().cpu().pin_memory() for k in keys}
        h2d = d2h = sum(v.numel() * v.element_size() for v in host_state.values())
        e0.record()
        for i in range(args.steps):
            if graphed is not None:           # static buffers: copy in place (h_half is the first half of the doubled list)
                for k in keys:
                    if k != "h_half":
                        cur[k].copy_(host_state[k], non_blocking=True)
                graphed.h_edge2[:Eh].copy_(host_state["h_half"], non_blocking=True)
                graphed.h_edge2[Eh:].copy_(graphed.h_edge2[:Eh])
            else:
                for k in keys:
                    st[k] = host_state[k].to(dev, non_blocking=True)
            step(args.warmup + i)
            for k in keys:
                host_state[k].copy_(cur[k], non_blocking=True)
            torch.cuda.current_stream().synchronize()       # the host needs the result before the next step
        e1.record()
    barrier()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = total_mols / (per_pass * float(t_e2e.item()) / args.steps * 1e-3)

    # ---- roofline of the dominant kernel, measured live with CUDA events on the launching stream ----
    roof = None
    if rank == 0:
        prof = engine.profile_kernels(lambda: eager_step(0), reps=3)
        pk = peaks()
        name, info = max(((k, v) for k, v in prof.items() if k in engine.KERNEL_MMA_FLOP_PER_EDGE),
                         key=lambda kv: kv[1]["ms_total"])
        avg_ms = info["ms_total"] / max(info["launches"], 1)
        executed = engine.KERNEL_MMA_FLOP_PER_EDGE[name] * E / (avg_ms * 1e-3) / 1e12
        as_written = engine.KERNEL_LOGICAL_FLOP_PER_EDGE.get(name, 0.0) * E / (avg_ms * 1e-3) / 1e12
        traffic, traffic_src = ncu_traffic()
        per_kernel = {k: {"ms_per_step": round(v["ms_total"] / 3, 4), "launches_per_step": v["launches"] // 3,
                          **({"tflops_executed_mma": round(engine.KERNEL_MMA_FLOP_PER_EDGE[k] * E * v["launches"] / (v["ms_total"] * 1e-3) / 1e12, 1),
                              "tflops_as_written": round(engine.KERNEL_LOGICAL_FLOP_PER_EDGE.get(k, 0.0) * E * v["launches"] / (v["ms_total"] * 1e-3) / 1e12, 1)}
                             if k in engine.KERNEL_MMA_FLOP_PER_EDGE and v["ms_total"] > 0 else {})}
                      for k, v in prof.items()}
        roof = {"kernel": name, "bound": "tensor", "achieved": executed, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                "frac": executed / pk["tf_sust"], "frac_executed_mma": executed / pk["tf_sust"],
                "achieved_as_written": as_written, "frac_as_written": as_written / pk["tf_sust"],
                "gemms_per_tile_KxN": engine.KERNEL_GEMMS[name], "mma_per_gemm_k_step": 3,
                "traffic": traffic.get(name), "peak_source": pk["src"] + " (sustained bf16, kernel timed inside a long step)",
                "traffic_source": traffic_src if traffic.get(name) else None,
                "avg_launch_ms": avg_ms,
                "note": "achieved = executed tcgen05 MMA FLOPs (GEMMs after exact per-node hoisting x 3 split-fp16 MMAs: hi*hi + lo*hi + "
                        "hi*lo) of one launch / its CUDA-event time; achieved_as_written = the reference's as-written GEMM FLOPs of the "
                        "same layer part (SURVEY.md 8d) / the same time",
                "hbm_algorithmic_gbs": (512.0 * E + 2080.0 * N) / (avg_ms * 1e-3) / 1e9,
                "share_of_step": info["ms_total"] / max(sum(v["ms_total"] for v in prof.values()), 1e-9),
                "per_kernel": per_kernel}

    # ---- end-of-run gather of the sampled molecules (the only collective of the path; outside the timed region) ----
    if dist is not None and not train:
        from moldiff_b200.sharding import gather_predictions
        last_p = eager_step(0)
        got = gather_predictions([last_p["pred_node"], last_p["pred_pos"], last_p["pred_halfedge"]],
                                 d["batch_node"], d["batch_halfedge"], dist, dst=0)
        if rank == 0:
            assert got["n_graphs"] == total_mols
        dist.barrier()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline and not train and not args.simple:
            threads = os.cpu_count() or 1
            cb = args.cpu_batch or (32 if workload == "guided" else 64)
            sec, cn, ceh = cpu_step_seconds(workload, cb, threads, reps=2)
            cpu = {"value": cb / (T_STEPS * sec), "unit": "molecules/s", "cores": threads, "kind": "port",
                   "sample": f"min of 2 loop bodies at B={cb} (N={cn}, E={2 * ceh}) after 1 warm-up; "
                             f"{sec:.2f} s/step; per-step cost is linear in E"}
        if train:
            metric = "molecules/sec through MolDiff.get_loss (forward + loss)"
            what = (f"train_MolDiff.yml forward + loss (BASELINE config 3): sample_time + add_noise + denoiser forward + losses, "
                    f"batch_size={B}/GPU")
        else:
            metric = f"molecules/sec at {T} denoise steps"
            what = ((f"sample_MolDiff_simple.yml (BASELINE config 1, T=50)" if args.simple else f"sample_MolDiff.yml {workload}")
                    + ": one loop body of the sampler (denoiser fwd + posterior sampling"
                    + (" + bond-predictor guidance fwd+bwd" if guidance else "") + f"), batch_size={B}/GPU"
                    + (", replayed as one CUDA graph" if graphed is not None else ""))
        line = {
            "metric": metric, "value": value, "unit": "molecules/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32 (GEMMs: split-fp16 hi|lo operands, 3 tcgen05 MMAs per product, fp32 TMEM accumulate; everything else fp32)",
            "data": "synthetic",
            "config": {"workload": what + ", " + (f"every molecule {args.max_size} atoms (QM9-sized dense batch)" if args.max_size
                                                  else "GEOM-Drugs node-count distribution"),
                       "global_batch": total_mols, "n_nodes_rank0": N, "n_edges_rank0": E, "parallelism": f"dp{world}",
                       "l2": "256 MiB flush write between timed iterations", "weights": "random-init (seed 0)",
                       "cuda_graph": graphed is not None,
                       "side_stream_overlap": os.environ.get("MDB_OVERLAP", "1") != "0"},
            "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
        }
        if per_rank_ms is not None:
            line["per_rank_ms_per_step"] = per_rank_ms
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
