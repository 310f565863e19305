#!/usr/bin/env python
"""G+D training-step throughput (BASELINE.json metric: "G+D train images/sec at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W [--config 3stages] [--impl reference]

A step = one pass of the hot path over one per-GPU batch of synthetic data: G forward, one update per
discriminator (real/wrong/fake forwards + backward + Adam), generator loss through the updated discriminators,
G backward + Adam, gradient all-reduce when N > 1.  The whole step is one CUDA-graph replay.
Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = same metric through the public
trainer API with host batches (pinned H2D copies and a D2H loss read inside the timed region).
`--impl reference` times the reference's own CPU step (oracle port, or the real reference tree when present).
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

METRIC = "G+D train images/sec"
UNIT = "images/s"
# algorithmic GFLOP per image, reference dense-conv counting, step = 3*G_fwd + 12*sum(D_fwd)  
GFLOP_PER_IMAGE = {"catcls": 48.98, "3stages": 143.1, "onlycapsule": 50.20, "splitz_cap_ca": 54.37, "coco": 48.94}
WORKLOAD = {"catcls": "cfg/birds_2stgs_catcls.yml", "3stages": "cfg/birds_3stages.yml",
            "onlycapsule": "cfg/birds_2stgs_onlycapsule.yml",
            "splitz_cap_ca": "cfg/birds_2stg_splitz_cap_ca.realcls.yml", "coco": "cfg/coco_2stgs.yml"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.thr.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_step_rate(config, batch, steps, warmup, budget_s, threads=None):
    """Reference CPU step (fp32, torch CPU): the real reference tree when present, else the oracle port.
    Returns (images/s, kind, cores, sample description)."""
    import torch
    from oracle import configs as ocfg, ref_harness, shapes, synth
    from oracle.ekl_oracle import OracleTrainer
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    oc = ocfg.oracle_cfg(config, batch=batch)
    if ref_harness.available():
        kind = "reference"
        torch.manual_seed(0)
        _, netG, netsD = ref_harness.build_nets(config, batch=batch)
        for n in [netG] + netsD:
            n.apply(ref_harness.import_reference()["cub"].weights_init)
        stepper = ref_harness.RefStepper(config, netG, netsD)
    else:
        kind = "port"
        gsh = shapes.g_shapes(oc, cond_dim=ocfg.cond_dim(oc))
        dsh = [shapes.d_shapes(oc, r, True, oc.D_CAPSULE) for r in [64, 128, 256][: oc.BRANCH_NUM]]
        stepper = OracleTrainer(oc, shapes.make_state_dict(gsh, "G"),
                                [shapes.make_state_dict(s, "D%d" % i) for i, s in enumerate(dsh)])
    b = synth.make_batch(oc, batch, "bench")
    t_start = time.time()
    times = []
    for i in range(warmup + steps):
        t0 = time.time()
        stepper.step(**b)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_start > budget_s and times:
            break
    t = sum(times) / len(times)
    return batch / t, kind, cores, "%d timed step(s) of batch %d (full reference step, fp32 torch CPU), %.2f s/step" % (len(times), batch, t), t


def gpu_eager_rate(config, batch, steps=3, warmup=2):
    """The reference's step as plain PyTorch eager ops on this GPU (cuDNN / cuBLAS kernels, fp32 NCHW): the oracle port
    with every tensor on the device -- what a user of the unmodified reference gets from a B200 ("the real bar
    to beat").  'stock' = torch's default precision flags (TF32 allowed for cuDNN convolutions, fp32 matmuls);
    'strict_fp32' = TF32 off everywhere.  A reported baseline like cpu_baseline, never part of the product path."""
    import torch
    from oracle import configs as ocfg, shapes, synth
    from oracle.ekl_oracle import OracleTrainer
    dev = torch.device("cuda", torch.cuda.current_device())
    oc = ocfg.oracle_cfg(config, batch=batch)
    to_dev = lambda v: [t.to(dev) for t in v] if isinstance(v, (list, tuple)) else (v.to(dev) if torch.is_tensor(v) else v)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    res = {}
    try:
        for mode, conv_tf32 in (("stock", True), ("strict_fp32", False)):
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = conv_tf32, False
            gsh = shapes.g_shapes(oc, cond_dim=ocfg.cond_dim(oc))
            dsh = [shapes.d_shapes(oc, r, True, oc.D_CAPSULE) for r in [64, 128, 256][: oc.BRANCH_NUM]]
            on_dev = lambda sd: {k: v.to(dev) for k, v in sd.items()}
            stepper = OracleTrainer(oc, on_dev(shapes.make_state_dict(gsh, "G")),
                                    [on_dev(shapes.make_state_dict(s, "D%d" % i)) for i, s in enumerate(dsh)])
            b = {k: to_dev(v) for k, v in synth.make_batch(oc, batch, "bench").items()}
            for _ in range(warmup):
                stepper.step(**b)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                stepper.step(**b)
            e1.record()
            torch.cuda.synchronize()
            res[mode] = batch / (e0.elapsed_time(e1) / steps * 1e-3)
            del stepper, b
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return {"unit": UNIT, "kind": "port, torch eager on cuda:0 (cuDNN / cuBLAS, fp32 NCHW)", "batch": batch,
            "sample": "%d timed steps after %d warm-up" % (steps, warmup), **res}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = a.batch or DEFAULT_BATCH[a.config]
    # bounded sample: shrink the batch until (steps + warmup) steps fit the budget (BatchNorm keeps work/image ~constant)
    budget = float(os.environ.get("EKL_REF_BUDGET_S", "240"))
    import torch
    probe_b = min(batch, 4)
    rate, kind, cores, sample, t = cpu_step_rate(a.config, probe_b, 1, 0, 1e9)
    per_img = t / probe_b
    b = batch
    while b > 2 and per_img * b * (a.steps + a.warmup) > budget:
        b //= 2
    rate, kind, cores, sample, t = cpu_step_rate(a.config, b, a.steps, a.warmup, budget)
    line = {"metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD[a.config], "resolved_config": a.config, "batch_per_step": b,
                       "note": "reference CPU training step on host cores; rank 0 only"},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


DEFAULT_BATCH = {"catcls": 24, "3stages": 24, "onlycapsule": 32, "splitz_cap_ca": 32, "coco": 64}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from text2img_ekl_b200 import _lib, configs, ops, parallel
    from text2img_ekl_b200.engine import GraphedStep
    from text2img_ekl_b200.synthetic import SyntheticLoader
    rank, ws = parallel.init_from_env()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().ekl_require_sm100())
    torch.backends.cuda.matmul.allow_tf32 = False
    B = a.batch or DEFAULT_BATCH[a.config]
    Trainer = configs.setup(a.config, batch=B)
    torch.manual_seed(0)
    tr = Trainer(None, None, 64)
    tr.setup()
    parallel.broadcast_params([tr.netG] + tr.netsD)
    cls_kind = getattr(tr, "CLS_KIND", "index")
    loader = SyntheticLoader(B, cls_kind, rank=rank, pool=4)
    pool = loader.pool
    use_graph = not a.no_graph
    if use_graph:
        gs = GraphedStep(tr, pool[0])
        launches_per_step = gs.launches_per_step
        step_resident = gs.replay
        step_e2e = lambda d, nxt=None: gs.step(d, nxt)
        h2d = gs.h2d_bytes
    else:
        n0 = ops.LAUNCHES[0]
        tr.train_step(pool[0])
        launches_per_step = ops.LAUNCHES[0] - n0
        dev_batches = [tuple([[t.to(tr.device) for t in x] if isinstance(x, list) else (x.to(tr.device) if x is not None else None)
                              for x in b]) for b in pool]
        step_resident = lambda: tr.train_step(dev_batches[0])
        step_e2e = lambda d, nxt=None: tr.train_step(d)
        h2d = sum(t.numel() * t.element_size() for t in pool[0][0] + pool[0][1] + [pool[0][2], pool[0][3]])

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if ws > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    # ---- device-resident throughput (inputs already in HBM)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(lambda i: step_resident(), a.steps, a.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms / a.steps
    value = ws * B / (ms_per_step * 1e-3)

    # ---- end to end through the public API: pinned host batch -> H2D -> step -> D2H loss read, every step
    host_loss = torch.zeros(8, pin_memory=True)

    def e2e_step(i):
        out = step_e2e(pool[i % len(pool)], pool[(i + 1) % len(pool)])   # upload of step i+1 overlaps step i
        errG = out[1] if isinstance(out, tuple) else out
        v = errG if torch.is_tensor(errG) else torch.stack([x.detach().float() for x in errG])
        host_loss[: v.numel()].copy_(v.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the user reads the loss: a real D2H dependency every step
    ms2 = timed(e2e_step, a.steps, max(a.warmup, 1))
    e2e_value = ws * B / (ms2 / a.steps * 1e-3)
    d2h = 4 * 6

    line = None
    if rank == 0:
        pk = peaks()
        gf = GFLOP_PER_IMAGE[a.config]
        roof = kernel_roofline(tr, pool[0], pk) if not a.no_profile else None
        step_tf = gf * (value / ws) / 1e3
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ws, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": WORKLOAD[a.config], "resolved_config": a.config, "batch_per_gpu": B,
                           "global_batch": ws * B, "parallelism": "dp%d" % ws, "cuda_graph": use_graph,
                           "l2": "working set per step (activations + weights, > 1 GB) exceeds the 126 MB L2; no flush",
                           "gflop_per_image_reference_count": gf},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms2 / a.steps,
                        "pipeline": "one upload per step; batch k+1 goes up on a copy stream behind step k" if use_graph
                        else "upload on the critical path"},
                "gpu_launches": launches_per_step * a.steps,
                "step_tflops_reference_count": step_tf,
                "step_frac_of_bf16_sustained": step_tf / pk["bf16_sustained"],
                "roofline": roof}
    if rank == 0 and ws == 1 and not a.no_cpu:
        try:
            cb = min(B, int(os.environ.get("EKL_CPU_BASELINE_BATCH", "8")))
            rate, kind, cores, sample, _ = cpu_step_rate(a.config, cb, 1, 1, float(os.environ.get("EKL_CPU_BUDGET_S", "40")))
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %s" % ex}
    if rank == 0 and ws == 1 and not a.no_cpu:
        try:
            line["gpu_eager_baseline"] = gpu_eager_rate(a.config, B)
        except Exception as ex:  # noqa: BLE001
            line["gpu_eager_baseline"] = {"unit": UNIT, "failed": str(ex)[:200]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if ws > 1:
        # every rank is past its timed region; leave without tearing the NCCL communicator down under live CUDA
        # graphs that captured its kernels (destroy_process_group was seen to hang there)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def ncu_traffic(family):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture (profiles/r01_conv_full.json, written by tools/ncu_summary.py) -> (bytes per launch, provenance);
    (None, None) when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "r01_conv_full.json")
    if not os.path.exists(p):
        return None, None
    key = "conv_wgrad" if "wgrad" in family else "conv_gemm_tc"
    rows = [r for r in json.load(open(p)) if key in r.get("kernel", "")]
    if not rows:
        return None, None
    tot = sum(r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0) for r in rows)
    return tot / len(rows), {"launches_captured": len(rows), "source": "profiles/r01_conv_full.json"}


def kernel_roofline(tr, batch, pk):
    """One eager step with a CUDA-event pair around every kernel-library call (on the launching stream), aggregated
    per kernel family; `roofline` describes the dominant family (tcgen05 conv forward/dgrad kernel)."""
    import torch
    from text2img_ekl_b200 import ops
    ops.PROFILE = []
    # Park the GPU behind a ~120 ms spin before every phase of the step (generate, each discriminator update, generator
    # update) so that the host enqueues the whole phase ahead of the device (each phase stays below the launch-queue
    # depth): every event pair then brackets back-to-back device execution of its kernel(s), not host launch latency.
    spin = int(0.12 * 1.9e9)
    torch.cuda.synchronize()
    tr.imgs_tcpu, tr.real_imgs, tr.wrong_imgs, tr.txt_embedding, tr.cls_label = tr.prepare_data(batch)
    tr.noise.normal_(0, 1)
    torch.cuda._sleep(spin)
    tr.generate()
    for i in reversed(range(tr.num_Ds)):
        torch.cuda.synchronize()
        torch.cuda._sleep(spin)
        tr.train_joint_Dnet(i, 1)
    torch.cuda.synchronize()
    torch.cuda._sleep(spin)
    tr.engine.g_step(tr.real_cp)
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    fam = {}
    for name, flops, nbytes, e0, e1 in rec:
        d = fam.setdefault(name, dict(us=0.0, flop=0.0, bytes=0.0, launches=0))
        d["us"] += e0.elapsed_time(e1) * 1e3
        d["flop"] += flops
        d["bytes"] += nbytes
        d["launches"] += 1
    total = sum(d["us"] for d in fam.values()) or 1.0
    out = {}
    for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        out[k] = {"us": round(d["us"], 1), "share": round(d["us"] / total, 4), "launches": d["launches"],
                  "tflops": round(d["flop"] / (d["us"] * 1e-6) / 1e12, 1) if d["flop"] else None,
                  "gbs": round(d["bytes"] / (d["us"] * 1e-6) / 1e9, 1) if d["bytes"] else None}
    top = max((k for k in fam if fam[k]["flop"] > 0), key=lambda k: fam[k]["us"], default=None)
    if top is None:
        return None
    d = fam[top]
    ach = d["flop"] / (d["us"] * 1e-6) / 1e12
    traffic, traffic_src = ncu_traffic(top)
    return {"bound": "tensor", "kernel": top, "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
            "frac": ach / pk["bf16_sustained"], "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["src"] + " (sustained: kernel timed inside a long step)",
            "flop_counting": "algorithmic (reference dense-conv count) flops of the launches / sum of their CUDA-event durations",
            "avg_launch_us": d["us"] / d["launches"], "launches_per_step": d["launches"], "families": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="3stages", choices=sorted(WORKLOAD))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
        return
    if a.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(a)


if __name__ == "__main__":
    main()
