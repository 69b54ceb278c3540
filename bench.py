#!/usr/bin/env python
"""bench.py — self-play throughput of the B200 engine on BASELINE.json's headline config.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload = "C3"): 19x19 Go (wq), 1024 concurrent games per GPU, 800 sims/move,
20-block x 256 dual net (F=18, FC=512, A'=362, DefaultConf batch 256), two random-init nets (agents A
and B, agogo.go:50-58), synthetic data = the self-play positions themselves.
  step   = one MCTS wave over all resident games: PUCT select/expand-prep for every game ->
           batched dual-net evaluation of both agents' leaves -> expand + backup
           (i.e. n_games simulations; every 801st wave is a move's root evaluation).
  value  = simulations/s over all GPUs, device-timed (CUDA events inside the engine are used for
           the kernel roofline; the step loop itself is bracketed by barrier + synchronize).
  e2e    = the same metric through the reference-facing call az_arena_step (one full Arena.Play
           ply: Search + Apply + Example read-back into host buffers), wall clock, host<->device
           copies inside the timed region.
The reference arm times the CPU oracle (restatement of the reference algorithm; no Go toolchain in
this image) on the host cores: one step = one simulation of one game.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from agogo_b200 import _capi as K  # noqa: E402

WORKLOADS = {
    # name: (size, n_games, sims, k, blocks, fc, batch)
    "C3": dict(size=19, n_games=1024, sims=800, k=256, blocks=20, fc=512, batch=256),
    "C2": dict(size=9, n_games=256, sims=400, k=64, blocks=6, fc=128, batch=256),
}


def flops_per_eval(w):
    hw = w["size"] ** 2
    k, b, fc, a1 = w["k"], w["blocks"], w["fc"], hw + 1
    tower = 2 * b * (2 * 9 * k * k * hw) + 2 * 9 * 18 * k * hw
    heads = 2 * k * 2 * hw + 2 * 2 * hw * a1 + 2 * k * hw + 2 * hw * fc + 2 * fc
    return tower + heads


def make_desc(w, n_games, device, seed, sims=None, batch=None):
    s = w["size"]
    return K.make_desc(K.GAME_WQ, s, s, 0, komi=7.5, sims=sims if sims is not None else w["sims"], n_games=n_games,
                       seed=seed, device=device, max_moves=2 * s * s,
                       nn=dict(k=w["k"], shared_layers=w["blocks"], fc=w["fc"], batch_size=batch or w["batch"],
                               features=18, action_space=s * s + 1))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]),
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows), "reasons": reasons}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def oracle_lib():
    so = os.path.join(ROOT, "oracle", "libazoracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return K.load(so)


def setup_nets(e, seed):
    e.net_init(0, seed + 100)
    e.net_init(1, seed + 101)
    e.set_inferer(0, K.INF_DUAL)
    e.set_inferer(1, K.INF_DUAL)


def cpu_reference_run(w, steps, warmup, threads=None):
    """One step = one pipeline() iteration (1 leaf evaluation at useful-work batch 1) of ONE game on the
    oracle, OpenMP over the host cores inside the conv loops."""
    lib = oracle_lib()
    ncores = len(os.sched_getaffinity(0))
    # small train batch: the oracle only reads batch row 0 at inference; keeps its memory modest
    e = lib.create(make_desc(w, 1, 0, 1234, batch=2))
    setup_nets(e, 1234)
    e.arena_begin(1, False)
    e.search_begin()  # root evaluation
    # "all the host threads it can use": try the full set and a few smaller teams (shared hosts / cgroup quotas
    # make the largest team the slowest one surprisingly often) and keep the fastest
    set_threads = lib.dll.azo_set_threads
    best_t, best_dt = None, None
    for t in sorted({ncores, max(1, ncores // 2), max(1, ncores // 4), max(1, ncores // 8)}, reverse=True):
        set_threads(t)
        e.search_run(1)
        t0 = time.perf_counter()
        e.search_run(1)
        d = time.perf_counter() - t0
        if best_dt is None or d < best_dt:
            best_t, best_dt = t, d
    set_threads(best_t)
    for _ in range(warmup):
        e.search_run(1)
    t0 = time.perf_counter()
    e.search_run(steps)
    dt = time.perf_counter() - t0
    c = e.counters()
    e.close()
    return dict(value=steps / dt, seconds=dt, cores=best_t, evals=c["evals"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=list(WORKLOADS))
    ap.add_argument("--games", type=int, default=0, help="override games per GPU (debug)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=24)
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.games:
        w["n_games"] = args.games
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)
    fpe = flops_per_eval(w)
    config = {"workload": "%s: %dx%d Go (wq) self-play, %d games/GPU, %d sims/move, %d-block x %d dual net, two random-init "
                          "nets" % (args.workload, w["size"], w["size"], w["n_games"], w["sims"], w["blocks"], w["k"]),
              "games_per_gpu": w["n_games"], "sims_per_move": w["sims"],
              "parallelism": "games sharded across %d GPU(s), no collective on the self-play path" % world,
              "step": "one MCTS wave = n_games simulations (select -> batched dual-net eval -> expand/backup)",
              "l2": "inputs larger than L2: activations ~%d MB per conv layer per agent vs 126 MB L2" %
                    (w["n_games"] // 2 * (w["size"] + 1) ** 2 * w["k"] * 4 * 2 // 2 ** 20),
              "flops_per_eval": fpe}

    if args.impl == "reference":
        if rank != 0:
            return 0
        # torchrun pins OMP_NUM_THREADS=1; the reference arm gets every host thread this process may run on
        # (set before libgomp loads); passive waiting keeps it sane on shared hosts
        if os.environ.get("OMP_NUM_THREADS", "1") == "1":
            os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        r = cpu_reference_run(w, args.steps, warmup)
        line = {"impl": "reference", "metric": "mcts_sims_per_sec", "value": r["value"], "unit": "sims/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(config, step="one simulation of one game (bounded sample of the GPU arm's wave)"),
                "cpu_baseline": {"value": r["value"], "unit": "sims/s", "cores": r["cores"], "kind": "port",
                                 "sample": "oracle (C++ restatement of mcts+dualnet), 1 game, %d pipeline iterations, "
                                           "1 leaf eval each at useful-work batch 1 (the reference pads every eval to "
                                           "ActionSpace=%d samples, meta.go:125-135: divide by that for its faithful rate)"
                                           % (args.steps, w["size"] ** 2)},
                "e2e": {"value": r["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "moves_per_sec": r["value"] / (w["sims"] + 1)}
        print(json.dumps(line))
        return 0

    dist = torch = None
    if world > 1:  # plumbing only: process group for barrier / max-over-ranks
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # the VERSION banner goes to stdout; stdout carries exactly one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = K.load()
    e = lib.create(make_desc(w, w["n_games"], local_rank, 1000 + rank))
    setup_nets(e, 1000 + 10 * rank)
    n_games = w["n_games"]

    def barrier():
        e.counters()  # synchronises the engine's stream (all of this process's GPU work)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(local_rank)

    # step generator: waves continue across move boundaries
    state = {"in_search": False, "left": 0}
    e.arena_begin(n_games, False)

    def one_step():
        if not state["in_search"]:
            e.search_begin()          # root evaluation wave (1 eval per game)
            state["in_search"], state["left"] = True, w["sims"]
            return 0
        e.search_run(1)
        state["left"] -= 1
        if state["left"] == 0:
            e.search_end()
            state["in_search"] = False
        return 1

    for _ in range(warmup):
        one_step()

    def timed_region():
        e.counters_reset()
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e.profile(True)                      # records the start event on the engine's stream
        for _ in range(args.steps):
            one_step()
        prof = e.profile(False)              # stop event + synchronise: device time of exactly `steps` steps
        barrier()
        sampler.stop_flag = True
        return prof, sampler.summary(), e.counters()

    prof, clocks, cnt = timed_region()
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if bad & set(clocks.get("reasons", [])) and world == 1:  # rejected: re-measure once (timing rules)
        clocks["rejected_first_run"] = True
        prof, clocks2, cnt = timed_region()
        clocks2["rejected_first_run"] = True
        clocks = clocks2
    dt = prof["region_ms"] / 1e3
    sims = cnt["sims"]
    evals = cnt["evals"]
    tot = [dt, float(sims), float(evals), float(cnt["kernel_launches"])]
    dt_max = dt
    if dist is not None:
        tt = torch.tensor(tot, dtype=torch.float64, device="cuda:%d" % local_rank)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dt_max, tot = mx[0].item(), tt.tolist()
    tot_sims, tot_evals, tot_launch = tot[1], tot[2], tot[3]
    value = tot_sims / dt_max

    # ---- e2e: one full Arena.Play ply through az_arena_step, examples read back to host
    e2e = None
    if not args.no_e2e:
        if state["in_search"]:  # finish the ply in flight (untimed)
            e.search_run(state["left"])
            e.search_end()
            state["in_search"] = False
        e.arena_finish()
        e.examples(clear=True)
        e.arena_begin(n_games, True)
        barrier()
        t1 = time.perf_counter()
        e.arena_step()
        barrier()
        dt_e = time.perf_counter() - t1
        if dist is not None:
            te = torch.tensor([dt_e], dtype=torch.float64, device="cuda:%d" % local_rank)
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dt_e = te[0].item()
        d2h = n_games * (18 * w["size"] ** 2 + w["size"] ** 2 + 1 + 2) * 4 + 8
        e2e = {"value": world * n_games * w["sims"] / dt_e, "unit": "sims/s", "h2d_bytes_per_step": n_games * 4,
               "d2h_bytes_per_step": d2h, "step": "az_arena_step: one full ply (1 root eval + %d sims per game) + example read-back"
               % w["sims"], "seconds": dt_e, "moves_per_sec": world * n_games / dt_e,
               "note": "the path's inputs are the two nets' weights, uploaded once per epoch by az_net_set_params / "
                       "az_agent_set_inferer, not per ply; per ply the host sends the coin flips and receives the examples"}

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peaks, peak_src = load_peaks()
    hw = w["size"] ** 2
    # dominant kernel: the fused 3x3 conv of one residual block (both branches, C -> 2C), one launch per
    # block per agent.  Algorithmic FLOPs of all its launches in the timed region:
    conv_flops_per_eval = 2 * 9 * w["k"] * (2 * w["k"]) * hw
    # every evaluated leaf passes through `blocks` such launches; evals on this rank:
    conv_flops = conv_flops_per_eval * evals * w["blocks"]
    conv_s = prof["conv_ms"] / 1e3
    achieved = conv_flops / conv_s / 1e12 if conv_s > 0 else None
    peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_conv2_traffic.json")
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "tensor", "kernel": "k_conv3x3_tc2 (fused residual-block conv, CTA pair / cta_group::2, fp16 hi/lo 3-pass tcgen05)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": "%s bf16 dense, sustained (kernel timed inside a long step)" % peak_src,
                "traffic": traffic, "launches_timed": prof["conv_launches"],
                "avg_launch_ms": prof["conv_ms"] / max(prof["conv_launches"], 1),
                "algorithmic_flops_per_launch": conv_flops / max(prof["conv_launches"], 1),
                "share_of_step": conv_s / dt, "note": "3 tensor-core passes per algorithmic MAC (fp32-faithful split): "
                "frac of the bf16 peak tops out at 1/3 x 361/384 (M-tile padding per sample) = 0.31"}
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")
        r = cpu_reference_run(w, args.cpu_steps, 1)
        cpu_baseline = {"value": r["value"], "unit": "sims/s", "cores": r["cores"], "kind": "port",
                        "sample": "oracle, 1 game x %d pipeline iterations (1 leaf eval each, useful-work batch 1), %.1f s"
                                  % (args.cpu_steps, r["seconds"])}
    # tree-kernel traffic from COUNTED events: 12 B per child scanned by Select (N, W, P), 8 B per level
    # (meta, first), 20 B per node created, 8 B read-modify-write per node backed up
    tree_bytes = 12 * cnt["select_children"] + 8 * cnt["select_levels"] + 20 * cnt["created"] + 8 * cnt["backup_nodes"]
    tree = {"counted_bytes_per_sim": tree_bytes / max(sims, 1), "select_children_per_sim": cnt["select_children"] / max(sims, 1),
            "levels_per_sim": cnt["select_levels"] / max(sims, 1), "nodes_created_per_eval": cnt["created"] / max(evals, 1),
            "null_results": cnt["null_results"],
            "note": "k_select + k_expand_backup take ~0.14 ms of a ~45 ms wave (profiles/r01_summary.md): latency-bound "
                    "pointer chase, not bandwidth-bound"}
    line = {"metric": "mcts_sims_per_sec", "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": 1e3 * dt_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32",
            "precision": "fp32 semantics on fp16 tensor cores: hi/lo operand split, 3 tcgen05 passes, fp32 TMEM accumulate; "
                         "policy/value within 1.4e-5 of the fp32 oracle (tolerance 1e-4)",
            "data": "synthetic", "config": config, "moves_per_sec": value / (w["sims"] + 1),
            "evals_per_sec": tot_evals / dt_max, "tflops_algorithmic": tot_evals / dt_max * fpe / 1e12,
            "e2e": e2e, "gpu_launches": int(tot_launch), "roofline": roofline, "tree": tree, "cpu_baseline": cpu_baseline,
            "clocks": clocks, "timing": "CUDA events on the engine's stream around exactly `steps` steps, barrier+synchronize both sides, max over ranks"}
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
