#!/usr/bin/env python
"""bench.py — AVBD step-loop throughput on B200 (BASELINE.json metric: steps/s at Stress1000 & 1M-box,
body-solves/s/GPU, HBM GB/s vs peak).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload grid100|stress1000|ensemble] [--impl reference]

One "step" = one Solver::step() (broadphase -> narrowphase/warm-start -> predict -> iterations x (primal per
colour, dual) -> velocities) over one resident world.  Default workload (N=1): the 100x100x100 = 1M-box
pre-stacked grid of SURVEY.md section 8d(4), iterations=10 — the largest single-GPU configuration BASELINE.json
names and the one the HBM roofline is meaningful on; Stress1000 (the reference's own scene, latency-bound on a
GPU) is measured in the same run and reported under "stress1000".  N>1: one such world per GPU (independent
worlds, no data-path collective; NCCL only gathers diagnostics) => weak scaling.

Printed: ONE JSON line (rank 0).  `value` = body-solves/s of the whole job with state resident in HBM, timed
with CUDA events on the solver's stream, max over ranks.  `e2e` = the same metric through the public API with
host buffers: every step uploads the body state from pinned host memory, steps, and downloads it back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

PRIMAL_BYTES_PER_BODY, PRIMAL_BYTES_PER_VISIT, DUAL_BYTES_PER_CONTACT = 100, 124, 156      # SURVEY.md section 8d / DESIGN.md
# A dual pass applied INSIDE a primal sweep (deferred dual) re-uses what the visit already loaded; the only
# compulsory traffic it adds is the penalty + stick write-back.  Counted this way the fused sweep gets no credit for the
# 156 B/contact pass it made unnecessary — removing traffic must not raise the "achieved" figure.
DEFERRED_DUAL_BYTES_PER_CONTACT = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="grid100", choices=["grid100", "grid50", "stress1000", "ensemble"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per primal colour sweep from the committed `ncu --set full` capture of this workload (profiles/r01_traffic.json,
    written by profiles/summarize.py traffic); None when the file is absent.  Cold-cache, serialised launches: an upper bound."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_colour_sweep"]), "profiles/r01_traffic.json (%s)" % t.get("source", "ncu")
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def workload_preset(name):
    from avbd_demo3d_b200 import scenes
    if name == "grid100":
        s = scenes.stress_grid(100, 100, 100, spacing_y=1.01, start_y=0.51, wide_ground=True)
        s["params"]["iterations"] = 10
        return s, "synthetic 100x100x100 = 1M-box pre-stacked drop grid (Stress1000 generator, spacingY=1.01, startY=0.51, widened ground), iterations=10"
    if name == "grid50":
        s = scenes.stress_grid(50, 50, 50, spacing_y=1.01, start_y=0.51, wide_ground=True)
        s["params"]["iterations"] = 10
        return s, "synthetic 50x50x50 = 125k-box pre-stacked grid, iterations=10"
    if name == "stress1000":
        return scenes.scene("Stress1000"), "--scene Stress1000 (iterations=20 beta=30000 gamma=0.995), settled for 400 steps"
    if name == "ensemble":
        s = scenes.ensemble(scenes.scene("Pyramid"), 8192)
        return s, "batched ensemble of 8192 independent Pyramid worlds (56 bodies each), iterations=10"
    raise ValueError(name)


def cpu_reference_rate(sample_n, steps, kind):
    """body-solves/s of the CPU implementation on a bounded sample: an n^3 pre-stacked sub-grid of the workload."""
    from _libs import Oracle, ref_available
    from avbd_demo3d_b200 import scenes
    use_ref = kind == "reference" and ref_available()
    o = Oracle("ref" if use_ref else "port").create()
    s = scenes.stress_grid(sample_n, sample_n, sample_n, spacing_y=1.01, start_y=0.51, wide_ground=True)
    for i in range(len(s["size"])):
        o.add_body(s["size"][i], float(s["density"][i]), float(s["friction"][i]), s["pos"][i], s["quat"][i], s["lin"][i], s["ang"][i])
    o.set_params(iterations=10, beta=30000.0, gamma=0.995)
    o.step(2)
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    dyn = len(s["size"]) - 1
    o.close()
    return dict(value=dyn * 10 * steps / dt, unit="body-solves/s", cores=1, kind="reference" if use_ref else "port",
                sample=f"{sample_n}^3={dyn}-box pre-stacked sub-grid of the workload, iterations=10, {steps} steps in {dt:.1f}s "
                       f"(reference broadphase is O(n^2): the full 1M grid is ~2600 s/step extrapolated)",
                steps_per_s=steps / dt)


def cpu_stress1000_rate(steps=40):
    from _libs import Oracle, ref_available
    o = Oracle("ref" if ref_available() else "port").create()
    o.load_scene("Stress1000")
    o.step(400)        # settle into the contact-heavy regime the GPU number is quoted on
    t0 = time.perf_counter()
    o.step(steps)
    dt = time.perf_counter() - t0
    o.close()
    return steps / dt


def run_reference(args, rank):
    if rank != 0:
        return
    sample_n, steps = 16, max(1, args.steps)
    t0 = time.perf_counter()
    per = []
    for _ in range(1):
        per.append(cpu_reference_rate(sample_n, args.warmup + steps, "reference"))
    r = per[0]
    _, desc = workload_preset(args.workload if args.workload != "stress1000" else "grid100")
    line = dict(metric="body_solves_per_s", value=r["value"], unit="body-solves/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 / r["steps_per_s"], higher_is_better=True, scaling="strong" if args.workload == "ensemble" else "weak", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference", config=dict(workload=desc, sample=r["sample"]),
                cpu_baseline=dict(value=r["value"], unit=r["unit"], cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                e2e=dict(value=r["value"], unit="body-solves/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t0)
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import avbd_demo3d_b200 as avbd
    from avbd_demo3d_b200 import scenes

    dist = None
    if world_size > 1:
        # stdout carries ONE JSON line: keep NCCL's "NCCL version ..." banner (printed from level VERSION up, which some images
        # configure) out of it; an explicit INFO / TRACE request is left alone
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    preset, desc = workload_preset(args.workload)
    if args.workload == "ensemble" and world_size > 1:      # shard the worlds (block partition) across ranks
        total = 8192
        per = total // world_size
        preset = scenes.ensemble(scenes.scene("Pyramid"), per, first_world=rank * per)
    w = avbd.World(local_rank)
    scenes.load(w, preset)
    n_bodies = w.n
    # setup (not warm-up): Stress1000 settles into its contact-heavy regime; the grids run until the manifold count
    # has plateaued so no device buffer grows inside the timed region
    w.step(400 if args.workload == "stress1000" else 12)
    iters = w.params["iterations"]
    w.step(max(3, args.warmup))
    stats0 = w.step_stats()
    n_dyn = stats0["dynamicBodies"]

    # ---- timed region: K steps, state resident in HBM
    w.set_profiling(True)
    launches0 = w.profile()["kernel_launches"]
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms = w.step_timed(args.steps)
    barrier()
    prof = w.profile()
    w.set_profiling(False)
    launches_timed = prof["kernel_launches"] - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    stats = w.step_stats()
    diag = w.diagnostics()

    # ---- end to end: host buffers in, host buffers out, every step
    host = torch.empty((n_bodies, 13), dtype=torch.float32).pin_memory()
    host_np = host.numpy()
    w.download_state_into(host_np)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        w.set_state(host_np)            # H2D of the step's inputs (pinned)
        w.step(1, sync=False)
        w.download_state_into(host_np)  # D2H of the step's result (pinned), synchronises
    d = w.diagnostics()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    # ---- diagnostics gather over NCCL (the only inter-GPU traffic of the path)
    gathered = None
    if dist is not None:
        mine = torch.tensor([diag["maxPen"], diag["maxLin"], float(diag["contacts"]), float(diag["manifolds"])], device=dev)
        allv = [torch.zeros_like(mine) for _ in range(world_size)]
        dist.all_gather(allv, mine)
        gathered = [[float(x) for x in v.tolist()] for v in allv]

    if rank == 0:
        total_dyn = n_dyn * world_size
        value = total_dyn * iters * args.steps / (ms_max * 1e-3)
        peak, peak_src = measured_peak()
        traffic, traffic_src = ncu_traffic()
        primal_bytes = (PRIMAL_BYTES_PER_BODY * prof["primal_bodies"] + PRIMAL_BYTES_PER_VISIT * prof["primal_visits"]
                        + DEFERRED_DUAL_BYTES_PER_CONTACT * prof["deferred_dual_contacts"])
        dual_bytes = DUAL_BYTES_PER_CONTACT * prof["dual_contacts"]
        primal_gbs = primal_bytes / max(prof["ms_primal"], 1e-9) / 1e6
        dual_gbs = dual_bytes / max(prof["ms_dual"], 1e-9) / 1e6
        line = dict(
            metric="body_solves_per_s", value=value, unit="body-solves/s", n_gpus=world_size, steps=args.steps, warmup=max(3, args.warmup),
            ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="strong" if args.workload == "ensemble" else "weak", vs_baseline=None, dtype="f32", data="synthetic",
            config=dict(workload=desc, bodies_per_gpu=n_bodies, dynamic_bodies_per_gpu=n_dyn, manifolds=stats["manifolds"], contacts=stats["contacts"],
                        colours=stats["colours"], iterations=iters, parallelism=f"independent-worlds x{world_size}",
                        l2="inputs larger than L2 (body + contact state > 126 MB)" if n_bodies > 300000 else "state is L2-resident; steady-state stepping, no flush"),
            steps_per_s=args.steps / (ms_max * 1e-3),
            roofline=dict(bound="hbm", kernel="primal_visit_flat<128,4> + primal_solve_flat (one pair per colour = one colour sweep; sweeps 2.. also apply the deferred dual)", achieved=primal_gbs, peak=peak, unit="GB/s", frac=primal_gbs / peak,
                          traffic=traffic if args.workload == "grid100" else None, traffic_source=traffic_src if args.workload == "grid100" else None,
                          peak_source=peak_src, algorithmic_bytes_per_launch=primal_bytes / max(prof["primal_launches"], 1),
                          avg_launch_ms=prof["ms_primal"] / max(prof["primal_launches"], 1), share_of_step=prof["ms_primal"] / ms,
                          deferred_dual_passes_per_step=prof["deferred_dual_contacts"] / max(1, prof["steps"] * max(1, stats["contacts"])),
                          dual=dict(kernel="dual_contacts (stand-alone passes only: the step's last)", achieved=dual_gbs, frac=dual_gbs / peak, share_of_step=prof["ms_dual"] / ms,
                                    avg_launch_ms=prof["ms_dual"] / max(prof["dual_launches"], 1))),
            stage_ms=dict(broadphase=stats["ms_broadphase"], narrowphase=stats["ms_narrowphase"], predict=stats["ms_predict"], graph=stats["ms_graph"],
                          primal_dual=stats["ms_primal"], velocity_diag=stats["ms_velocity"], total=stats["ms_total"]),
            e2e=dict(value=total_dyn * iters * e2e_steps / e2e_s, unit="body-solves/s", h2d_bytes_per_step=n_bodies * 52, d2h_bytes_per_step=n_bodies * 52 + 48,
                     steps=e2e_steps, ms_per_step=1e3 * e2e_s / e2e_steps),
            gpu_launches=int(launches_timed), clocks=clocks, diagnostics=diag, nccl_gathered_diagnostics=gathered)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_rate(16, 25, "reference")
            if args.workload in ("grid100", "grid50"):
                # the metric's other half: Stress1000 steps/s, GPU and CPU, same regime (settled, contact-heavy)
                w2 = avbd.World(local_rank)
                scenes.load(w2, scenes.scene("Stress1000"))
                w2.step(400)
                ms2 = w2.step_timed(200)
                st2 = w2.step_stats()
                w2.close()
                line["stress1000"] = dict(steps_per_s=200 / (ms2 * 1e-3), ms_per_step=ms2 / 200, manifolds=st2["manifolds"], contacts=st2["contacts"],
                                          cpu_steps_per_s=cpu_stress1000_rate(), cpu_cores=1)
        print(json.dumps(line), flush=True)
    w.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
