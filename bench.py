#!/usr/bin/env python
"""Benchmark of the LSDM denoising hot path (BASELINE.json metric: denoising-steps/sec = batch x timesteps / s).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--mode strict|hoisted] [--batch B]

Workload (config.workload): BASELINE configs[1] -- "SDM full 1000-step DDPM p_sample_loop, batch=64, 1xB200": T=1000
cosine schedule, B=64 samples per GPU, 9 clouds x 1024 points per sample, synthetic seeded inputs / random-init
(well-conditioned) weights.  One bench "step" = one denoising step of the whole batch (one pass of the hot path:
condition encode incl. PointNet++ on 9B clouds, x0 network, posterior + ancestral noise), K consecutive timesteps
999, 998, ...  STRICT mode (default) re-encodes the conditions every step from fresh FPS start draws, exactly the
per-step work of the reference; ``--mode hoisted`` encodes once (an algorithmic optimisation, reported separately).

value : device-timed (CUDA events on the launching stream), all inputs already resident in HBM.
e2e   : same metric through the reference-shaped public API (p_sample_loop over the same K timesteps) with HOST
        (pinned) condition buffers: host->device copies of conditions and per-step FPS starts and the final
        device->host read of the samples are inside the timed region.
N > 1 : weak scaling, B samples per GPU, samples sharded across ranks (global mask replicated), one NCCL all-gather of
        the outputs at the end of the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _claim_stdout():
    """stdout must carry ONE JSON line, but native libraries print there too (NCCL's "NCCL version ..." banner at communicator
    creation).  Keep a private duplicate of fd 1 for the result and point fd 1 at stderr for everything else."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return keep


def _emit(fd, line):
    os.write(fd, (json.dumps(line) + "\n").encode())

METRIC = "denoising-steps/sec (batch x timesteps)"
UNIT = "sample-steps/s"
WORKLOAD = "SDM 1000-step DDPM p_sample_loop, batch=64 per GPU, 9x1024-pt clouds (BASELINE configs[1])"
# algorithmic work per sample-step
FLOP_STRICT = 14.51e9
FLOP_HOISTED = 0.185e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="strict", choices=["strict", "hoisted"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_from=None, t_to=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ts, r in self.rows:
            if (t_from is not None and ts < t_from) or (t_to is not None and ts > t_to):
                continue  # only samples taken while the GPU was running this benchmark's kernels
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's p_sample, "as written" (materialised point attention), all host threads
# ----------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(steps, warmup, micro_batch=4):
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lsdm_oracle as O
    from lsdm_b200 import synthetic as syn

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = syn.make_state_dict(0, "wellcond")
    tables = O.diffusion_tables(O.cosine_betas(1000))
    inp = syn.make_inputs(1234, micro_batch)
    fps, noise = syn.make_step_randoms(4321, micro_batch, steps + warmup)
    x = inp["x_T"].clone()
    times = []
    with torch.no_grad():
        for k in range(steps + warmup):
            t = torch.full((micro_batch,), 999 - k, dtype=torch.long)
            t0 = time.perf_counter()
            out = O.p_sample(sd, tables, x, inp["mask"], t, inp["given_objs"], inp["given_cats"], inp["text_emb"], list(fps[k]),
                             noise[k], as_written=True)
            x = out["sample"]
            times.append(time.perf_counter() - t0)
    timed = times[warmup:]
    sec = sum(timed)
    return {"value": micro_batch * len(timed) / sec, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle port of the reference p_sample (as written, fp32 torch CPU, {cores} threads): micro-batch {micro_batch} x "
                      f"{len(timed)} timed steps after {warmup} warm-up; full B=64 x 1000-step workload is a linear extrapolation",
            "sec_per_step": sec / len(timed)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 1))
    cb = cpu_port_rate(steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": cb["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "mode": "strict", "note": "CPU port; bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from lsdm_b200 import synthetic as syn
    from lsdm_b200.model.sdm import SceneDiffusionModel
    from lsdm_b200.util.model_util import create_gaussian_diffusion, get_default_diffusion, get_default_model_proxd

    out_fd = _claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, (world, args.gpus)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    B, K, W = args.batch, args.steps, args.warmup
    Bg = B * world
    off = rank * B
    hoisted = args.mode == "hoisted"
    T = 1000

    model = SceneDiffusionModel(**{**get_default_model_proxd(), "device": local_rank})
    model.load_state_dict(syn.make_state_dict(0, "wellcond"))
    model.eval()
    diff = create_gaussian_diffusion(get_default_diffusion())
    model.set_shard(Bg, off)

    # seeded synthetic inputs at GLOBAL shape (every rank generates the same tensors and slices its shard; the global
    # mask is replicated because the reference's mask scrambles index it by global sample -- SURVEY.md 8e)
    inp = syn.make_inputs(1234, Bg)
    fps_all, noise_all = syn.make_step_randoms(4321, Bg, W + K)
    sl = slice(off, off + B)
    fps_loc = fps_all.view(W + K, 4, Bg, 9)[:, :, sl].reshape(W + K, 4, B * 9).contiguous()
    host = {"mask": inp["mask"].pin_memory(), "given_objs": inp["given_objs"][sl].contiguous().pin_memory(),
            "given_cats": inp["given_cats"][sl].contiguous().pin_memory(), "text_emb": inp["text_emb"][sl].contiguous().pin_memory(),
            "x_T": inp["x_T"][sl].contiguous().pin_memory()}
    g = {k: v.to(dev) for k, v in host.items()}
    fps_dev = fps_loc.to(dev)
    noise_dev = noise_all[:, sl].contiguous().to(dev)

    eng = diff._engine(model, B, dev)
    x = g["x_T"].clone()
    gather_buf = torch.empty(world, B, 1024, 3, device=dev) if world > 1 else None

    def run_steps(first_k, n, xbuf):
        return eng.sample_loop(xbuf, g["text_emb"], g["given_objs"], g["given_cats"], g["mask"], fps_dev[first_k:first_k + n],
                               noise_dev[first_k:first_k + n], T - 1 - first_k, hoisted)

    # warm-up (untimed).  The clock sampler was started before the set-up (nvidia-smi needs ~1 s to emit its first line); only
    # samples between here and the end of the last measured leg -- the GPU runs this benchmark's kernels back to back in
    # that window: warm-up, timed region, profiled pass, e2e, hoisted -- enter the reported median.
    t_load0 = time.perf_counter()
    run_steps(0, W, x)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = eng.launch_count()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run_steps(W, K, x)
    if world > 1:
        dist.all_gather_into_tensor(gather_buf.view(-1), x.view(-1))  # the path's single exchange step
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = Bg * K / (ms * 1e-3)

    # ---- per-kernel-class shares + roofline numerator: separate profiled pass (CUDA events around every launch) ----
    xp = g["x_T"].clone()
    eng.profile_begin()
    run_steps(W, min(K, 3), xp)
    cls_ms, cls_n, gemm_flops = eng.profile_end()
    detail = eng.profile_report()
    prof_steps = min(K, 3)
    tot_ms = sum(cls_ms.values())
    peaks, peak_src = measured_peaks()
    gemm_ms = cls_ms["gemm"]
    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # fp32 CUDA-core GEMM today; the tensor roofline it is judged against is TF32 dense = half the measured bf16 rate
    tf32_peak = peaks["bf16_tflops_sustained"] / 2.0
    traffic, traffic_src = None, None
    try:  # dram bytes per launch of the same kernel class from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "r1_ncu_kernel_metrics.json")) as f:
            tc = json.load(f)["_tensor_class"]
        traffic, traffic_src = tc["avg_dram_bytes_per_launch"], "profiles/r1_ncu_kernel_metrics.json (" + tc["source"] + ")"
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "tensor-core dense layers (sa_fused*, gemm_ws/gemm_tc, fp1_tail)", "achieved": gemm_tflops,
                "peak": tf32_peak, "unit": "TFLOP/s", "frac": gemm_tflops / tf32_peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": f"{peak_src} bf16_tflops_sustained/2 (TF32 dense runs at half the bf16 rate)",
                "flops_per_launch_avg": gemm_flops / max(1, cls_n["gemm"]), "launches_per_step": cls_n["gemm"] / prof_steps,
                "avg_launch_ms": gemm_ms / max(1, cls_n["gemm"]), "share_of_step": gemm_ms / tot_ms if tot_ms else None,
                "kernels": [{"kernel": r["tag"], "launches_per_step": r["launches"] / min(K, 3), "ms_per_step": r["ms"] / min(K, 3),
                             "achieved": (r["flops"] / (r["ms"] * 1e-3) / 1e12) if r["ms"] > 0 else 0.0,
                             "frac": (r["flops"] / (r["ms"] * 1e-3) / 1e12 / tf32_peak) if r["ms"] > 0 else 0.0}
                            for r in sorted(detail, key=lambda r: -r["ms"])],
                "algorithmic_tflops_whole_step": (FLOP_HOISTED if hoisted else FLOP_STRICT) * Bg * K / (ms * 1e-3) / 1e12}
    shares = {k: (v / tot_ms if tot_ms else 0.0) for k, v in cls_ms.items()}
    # The non-tensor kernel classes against the HBM roofline: ALGORITHMIC bytes per step (what the stage must read
    # and write once, from the tensor shapes; C = 9 B clouds) / the class's measured time.  They are latency- / issue-bound
    # exact-fp32 scans, not bandwidth-bound: the fractions say so.
    Cc = 9 * B
    lvl_n, lvl_s = (1024, 1024, 256, 64), (1024, 256, 64, 16)          # source points / centroids per SA level
    sel_bytes = {
        "fps": Cc * (1024 * 12 + sum(s * (4 + 12) for s in lvl_s) + 4 * 8),                     # xyz in; idx + xyz of 4 levels out
        "ball_query": Cc * sum(n * 12 + s * 12 + s * 32 * 4 for n, s in zip(lvl_n, lvl_s)),      # points + centroids in; groups out
        "three_nn": Cc * sum(f * 12 + c * 12 + f * 3 * 8 for f, c in ((64, 16), (256, 64), (1024, 256), (1024, 1024))),
        "sa_gather": Cc * (64 * 256 * 4 + 16 * 32 * 4 + 16 * 32 * 256 * 4),                       # level 4: projected rows + groups in; rows out
        "fp_combine": Cc * sum(c * ch * 4 + f * ch * 4 * 2 + f * 3 * 8 for f, c, ch in ((64, 16, 256), (256, 64, 256))),
    }
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_classes = []
    for name, nbytes in sel_bytes.items():
        c_ms = cls_ms.get(name, 0.0) / prof_steps
        if c_ms > 0:
            gbs = nbytes / (c_ms * 1e-3) / 1e9
            hbm_classes.append({"class": name, "ms_per_step": c_ms, "algorithmic_bytes_per_step": int(nbytes), "achieved_gbs": gbs,
                                "frac_of_hbm_peak": gbs / hbm_peak})
    roofline["hbm_bound_classes"] = {"peak_gbs": hbm_peak, "peak_source": f"{peak_src} hbm_gbs", "classes": hbm_classes}

    # ---- e2e through the public API with host buffers ----
    e2e = None
    if not args.no_e2e:
        from lsdm_b200.diffusion import gaussian_diffusion as gdm

        torch.manual_seed(7)
        # W untimed warm-up steps through the SAME public call (first-use costs of the API path: device RNG initialisation,
        # lazy module loads of torch's own kernels, pinned staging buffers)
        # (same call shape as the timed one, so that the caching allocator and the pinned staging buffers are warm: max(W, K) steps)
        diff.p_sample_loop_fused(model, (B, 1024, 3), host["mask"], host["given_objs"], host["given_cats"], host["text_emb"],
                                 noise=None, clip_denoised=False, device=dev, skip_timesteps=T - max(W, K), hoisted=hoisted,
                                 chunk=K).cpu()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # the reference's call: diffusion.p_sample_loop(model, shape, mask, given_objs, given_cats, y, clip_denoised=False);
        # skip_timesteps leaves K timesteps; host tensors are uploaded inside (Engine._f32), result read back.
        # Three back-to-back calls, each timed on its own with a synchronize on both sides; the best one is reported (a host
        # hiccup of tens of ms -- scheduler, nvidia-smi polling -- otherwise lands in a 50-100 ms region) and all three are listed.
        e2e_runs = []
        for _ in range(3):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            out = diff.p_sample_loop_fused(model, (B, 1024, 3), host["mask"], host["given_objs"], host["given_cats"], host["text_emb"],
                                           noise=None, clip_denoised=False, device=dev, skip_timesteps=T - K, hoisted=hoisted, chunk=K)
            out_host = out.cpu()
            torch.cuda.synchronize()
            e2e_runs.append(time.perf_counter() - t0)
        if world > 1:
            tdt = torch.tensor(e2e_runs, device=dev, dtype=torch.float64)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)  # per call: the slowest rank
            e2e_runs = [float(v) for v in tdt.tolist()]
        dt = min(e2e_runs)
        cond_bytes = sum(host[k].numel() * 4 for k in ("mask", "given_objs", "given_cats", "text_emb"))
        fps_bytes = 4 * 9 * B * 8
        e2e = {"value": Bg * K / dt, "unit": UNIT, "h2d_bytes_per_step": int(cond_bytes / K + fps_bytes),
               "d2h_bytes_per_step": int(out_host.numel() * 4 / K),
               "note": "p_sample_loop_fused over the same K timesteps via the reference-shaped API; conditions (pinned host) uploaded once per "
                       "call, FPS starts drawn on the CPU generator and uploaded per chunk, noise drawn on the device generator "
                       "(as the reference does), final samples read back; best of 3 calls",
               "runs_sample_steps_per_s": [Bg * K / v for v in e2e_runs]}

    # ---- hoisted variant (conditions encoded once per loop; an algorithmic optimisation, reported separately) ----
    hoisted_info = None
    if not hoisted:
        xh = g["x_T"].clone()
        eng.sample_loop(xh, g["text_emb"], g["given_objs"], g["given_cats"], g["mask"], fps_dev[:1], noise_dev[:W], T - 1, True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        eng.sample_loop(xh, g["text_emb"], g["given_objs"], g["given_cats"], g["mask"], fps_dev[:1], noise_dev[W:W + K], T - 1 - W, True)
        h1.record()
        torch.cuda.synchronize()
        hms = h0.elapsed_time(h1)
        if world > 1:
            thm = torch.tensor([hms], device=dev)
            dist.all_reduce(thm, op=dist.ReduceOp.MAX)
            hms = float(thm.item())
        hoisted_info = {"value": Bg * K / (hms * 1e-3), "unit": UNIT, "ms_per_step": hms / K,
                        "note": "conditions (incl. PointNet++) encoded once per K-step call instead of every step: NOT the reference's per-step work"}

    # ---- transparency leg: the same K steps with the closed-form selections for uniform (zero-padded) clouds switched off ----
    full_scans = None
    if not hoisted:
        eng.set_option("select_uniform", 0)
        try:
            xf = g["x_T"].clone()
            run_steps(0, W, xf)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            run_steps(W, K, xf)
            f1.record()
            torch.cuda.synchronize()
            fms = f0.elapsed_time(f1)
            if world > 1:
                tfm = torch.tensor([fms], device=dev)
                dist.all_reduce(tfm, op=dist.ReduceOp.MAX)
                fms = float(tfm.item())
            full_scans = {"value": Bg * K / (fms * 1e-3), "unit": UNIT, "ms_per_step": fms / K,
                          "identical_output": bool(torch.equal(xf, x)),
                          "note": "lsdm_set_option('select_uniform', 0): FPS / 3-NN run their full scans on clouds whose points all "
                                  "coincide too; `value` uses the closed forms (bit-identical selections, DESIGN.md 5)"}
        finally:
            eng.set_option("select_uniform", 1)
    clk = clocks.stop(t_load0, time.perf_counter()) if rank == 0 else None
    if rank == 0:
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_port_rate(4, 1)
            cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32" if eng.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "mode": args.mode, "precision": eng.precision + " (TF32 tcgen05 encoder, 3xTF32 x0 network, fp32 accumulate; selection kernels exact fp32)" if eng.precision == "tf32" else eng.precision, "global_batch": Bg, "per_gpu_batch": B, "timesteps_timed": K,
                       "parallelism": f"dp{world} (samples sharded, no data-path collective; one all-gather of outputs)",
                       "l2": "per-step working set (GBs of intermediates over 9*B clouds) is far larger than the 126 MB L2; no flush needed",
                       "weights": "seeded well-conditioned random init (lsdm_b200.synthetic)"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "kernel_time_shares": shares,
            "cpu_baseline": cpu_baseline, "hoisted": hoisted_info, "uniform_cloud_full_scans": full_scans,
        }
        _emit(out_fd, line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
