#!/usr/bin/env python
"""G+D training-step throughput (BASELINE.json metric: "G+D train images/sec at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W [--config 3stages] [--impl reference]

A step = one pass of the hot path over one per-GPU batch of synthetic data: G forward, one update per
discriminator (real/wrong/fake forwards + backward + Adam), generator loss through the updated discriminators,
G backward + Adam, gradient all-reduce when N > 1.  The whole step is one CUDA-graph replay.
Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = same metric through the public
trainer API with host batches (pinned H2D copies and a D2H loss read inside the timed region).
`--impl reference` times the reference's own CPU step (oracle port, or the real reference tree when present).

Every rank executes exactly the same sequence of steps (timed regions, the roofline pass, the barriers): any step
issues gradient all-reduces when N > 1, so nothing that runs a step may be rank-conditional.
"""
import argparse
import gc
import json
import os
import re
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
DEFAULT_BATCH = {"catcls": 24, "3stages": 24, "onlycapsule": 32, "splitz_cap_ca": 32, "coco": 64}


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
def cpu_step_rate(config, batch, steps, warmup, budget_s, threads=None, min_steps=1):
    """Reference CPU step (fp32, torch CPU): the real reference tree when present, else the oracle port.
    Runs `warmup` untimed + up to `steps` timed steps of the FULL batch; stops early (never below min_steps timed
    steps) once budget_s is used up.  Returns (images/s, kind, cores, sample description, s/step, timed steps)."""
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
        if time.time() - t_start > budget_s and len(times) >= min_steps:
            break
    t = sum(times) / len(times)
    sample = "%d timed step(s) after %d warm-up of batch %d (full reference step, fp32 torch CPU, %d threads), %.2f s/step" % (
        len(times), min(warmup, i), batch, cores, t)
    return batch / t, kind, cores, sample, t, len(times)


def gpu_eager_rate(config, batch, steps=3, warmup=2):
    """The reference's step as plain PyTorch eager ops on this GPU (cuDNN / cuBLAS kernels): the oracle port with every
    tensor on the device -- what a user of the unmodified reference gets from a B200 ("the real bar to
    beat").  'stock' = fp32 NCHW with torch's default precision flags (TF32 allowed for cuDNN convolutions, fp32
    matmuls); 'strict_fp32' = TF32 off everywhere; 'bf16_autocast_channels_last' = torch.autocast(bfloat16) with
    filters and images stored channels_last (cuDNN's tensor-core NHWC kernels), the strongest eager configuration.
    A reported baseline like cpu_baseline, never part of the product path."""
    import torch
    from oracle import configs as ocfg, shapes, synth
    from oracle.ekl_oracle import OracleTrainer
    dev = torch.device("cuda", torch.cuda.current_device())
    oc = ocfg.oracle_cfg(config, batch=batch)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    res = {}
    try:
        for mode, conv_tf32, autocast in (("stock", True, False), ("strict_fp32", False, False),
                                          ("bf16_autocast_channels_last", True, True)):
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = conv_tf32, autocast
            torch.backends.cudnn.benchmark = True            # the reference sets it (cub:286)
            cl = (lambda t: t.contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t) if autocast else (lambda t: t)
            gsh = shapes.g_shapes(oc, cond_dim=ocfg.cond_dim(oc))
            dsh = [shapes.d_shapes(oc, r, True, oc.D_CAPSULE) for r in [64, 128, 256][: oc.BRANCH_NUM]]
            on_dev = lambda sd: {k: cl(v.to(dev)) for k, v in sd.items()}
            stepper = OracleTrainer(oc, on_dev(shapes.make_state_dict(gsh, "G")),
                                    [on_dev(shapes.make_state_dict(s, "D%d" % i)) for i, s in enumerate(dsh)])
            to_dev = lambda v: [cl(t.to(dev)) for t in v] if isinstance(v, (list, tuple)) else (v.to(dev) if torch.is_tensor(v) else v)
            b = {k: to_dev(v) for k, v in synth.make_batch(oc, batch, "bench").items()}

            def one():
                if autocast:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        stepper.step(**b)
                else:
                    stepper.step(**b)
            try:
                for _ in range(warmup):
                    one()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    one()
                e1.record()
                torch.cuda.synchronize()
                res[mode] = batch / (e0.elapsed_time(e1) / steps * 1e-3)
            except Exception as ex:  # noqa: BLE001
                res[mode + "_failed"] = str(ex)[:160]
            del stepper, b
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return {"unit": UNIT, "kind": "port, torch eager on cuda:0 (cuDNN / cuBLAS)", "batch": batch,
            "sample": "%d timed steps after %d warm-up" % (steps, warmup), **res}


def run_reference_arm(a):
    """The reference's own CPU implementation of the step on the host cores, on the SAME config and batch as the GPU
    arm.  Bounded: one untimed warm-up step, then as many of the requested timed steps as fit EKL_REF_BUDGET_S (never
    fewer than one; the batch is never shrunk)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = a.batch or DEFAULT_BATCH[a.config]
    budget = float(os.environ.get("EKL_REF_BUDGET_S", "200"))
    rate, kind, cores, sample, t, n = cpu_step_rate(a.config, batch, max(a.steps, 1), min(a.warmup, 1), budget)
    line = {"metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD[a.config], "resolved_config": a.config, "batch_per_gpu": batch,
                       "global_batch": batch, "steps_timed": n,
                       "note": "reference CPU training step on host cores at the GPU arm's batch; rank 0 only; bounded sample: "
                               "the requested steps are cut short when they exceed %.0f s" % budget},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def make_trainer(config, B):
    import torch
    from text2img_ekl_b200 import configs
    Trainer = configs.setup(config, batch=B)
    torch.manual_seed(0)
    tr = Trainer(None, None, 64)
    tr.setup()                      # initialises torch.distributed from the environment and broadcasts rank 0's weights
    return tr


def timed_region(fn, steps, warmup, ws, on_gpu=True):
    """W untimed + K timed calls of fn, bracketed by barrier + synchronize; CUDA-event time, max over ranks (ms).
    on_gpu=False (the CPU control-flow test): wall clock, gloo."""
    import torch
    import torch.distributed as dist

    def barrier():
        if ws > 1:
            dist.barrier()
        if on_gpu:
            torch.cuda.synchronize()
    for i in range(warmup):
        fn(i)
    barrier()
    if on_gpu:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    else:
        t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    if on_gpu:
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1) if on_gpu else (time.perf_counter() - t0) * 1e3
    if ws > 1:
        t = torch.tensor([ms], device="cuda" if on_gpu else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def drive(resident, e2e, roofline, a, ws, on_gpu=True, after_resident=None):
    """The part of the benchmark that runs training steps, in the order every rank must follow: device-resident timed
    region, end-to-end timed region, roofline pass.  A step issues gradient all-reduces when ws > 1, so this sequence is
    identical on all ranks; only printing is rank-conditional (tests/test_parallel_gloo.py drives it with a stub step
    under gloo).  Returns (ms resident, ms e2e, roofline)."""
    ms = timed_region(resident, a.steps, a.warmup, ws, on_gpu)
    if after_resident is not None:
        after_resident()               # host-only bookkeeping (the clock sampler stops here)
    ms2 = timed_region(e2e, a.steps, max(a.warmup, 1), ws, on_gpu)
    roof = roofline() if (roofline is not None and not a.no_profile) else None
    return ms, ms2, roof


def measure(tr, a, ws, rank, local_rank, B, pk=None):
    """The timed regions (and, with pk, the roofline pass) of one configured trainer -> dict."""
    import torch
    from text2img_ekl_b200 import ops
    from text2img_ekl_b200.engine import GraphedStep
    from text2img_ekl_b200.synthetic import SyntheticLoader
    loader = SyntheticLoader(B, getattr(tr, "CLS_KIND", "index"), rank=rank, pool=4)
    pool = loader.pool
    use_graph = not a.no_graph
    gs = None
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
    # ---- end to end through the public API: pinned host batch -> H2D -> step -> D2H loss read, every step
    host_loss = torch.zeros(8, pin_memory=True)

    def e2e_step(i):
        out = step_e2e(pool[i % len(pool)], pool[(i + 1) % len(pool)])   # upload of step i+1 overlaps step i
        errG = out[1] if isinstance(out, tuple) else out
        v = errG if torch.is_tensor(errG) else torch.stack([x.detach().float() for x in errG])
        host_loss[: v.numel()].copy_(v.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the user reads the loss: a real D2H dependency every step

    def roofline():
        try:
            return kernel_roofline(tr, gs, pool[0], pk)
        except Exception as ex:  # noqa: BLE001
            if ws > 1:
                raise                # ranks must not diverge silently
            return {"failed": str(ex)[:300]}
    # ---- device-resident throughput (inputs already in HBM), then e2e, then the roofline pass: same order on all ranks
    sampler, clocks = ClockSampler(local_rank), []
    if rank == 0:
        sampler.start()
    ms, ms2, roof = drive(lambda i: step_resident(), e2e_step, roofline if pk is not None else None, a, ws,
                          after_resident=lambda: clocks.append(sampler.stop() if rank == 0 else None))
    clocks = clocks[0]
    return dict(ms_per_step=ms / a.steps, value=ws * B / (ms / a.steps * 1e-3), e2e_ms=ms2 / a.steps,
                e2e_value=ws * B / (ms2 / a.steps * 1e-3), h2d=h2d, d2h=4 * 6, launches=launches_per_step, clocks=clocks,
                graph=gs, use_graph=use_graph, roofline=roof)


def teardown(ws, *objs):
    """Leave cleanly: captured graphs hold NCCL kernels, so they are destroyed (and the device drained) BEFORE the
    process group.  A watchdog ends the process if the communicator teardown does not return (seen once in round 1
    when the group was destroyed under live graphs); the result line has already been printed by then."""
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    sys.stderr.flush()
    if ws <= 1:
        return
    dist.barrier()
    torch.cuda.synchronize()
    for o in objs:
        if o is not None and getattr(o, "graph", None) is not None:
            o.graph.reset()
            o.graph = None
    gc.collect()
    torch.cuda.synchronize()
    killer = threading.Timer(20.0, lambda: os._exit(0))
    killer.daemon = True
    killer.start()
    dist.destroy_process_group()
    killer.cancel()


def run_ours(a):
    import torch
    from text2img_ekl_b200 import _lib, parallel
    rank, ws = parallel.init_from_env()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    _lib.check(_lib.lib().ekl_require_sm100())
    torch.backends.cuda.matmul.allow_tf32 = False
    B = a.batch or DEFAULT_BATCH[a.config]
    tr = make_trainer(a.config, B)
    pk = peaks()
    m = measure(tr, a, ws, rank, local_rank, B, pk)
    gf = GFLOP_PER_IMAGE[a.config]
    roof = m["roofline"]
    step_tf = gf * (m["value"] / ws) / 1e3
    line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": ws, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD[a.config], "resolved_config": a.config, "batch_per_gpu": B,
                       "global_batch": ws * B, "parallelism": "dp%d" % ws, "cuda_graph": m["use_graph"],
                       "grad_comm": os.environ.get("EKL_GRAD_COMM", "bf16") if ws > 1 else None,
                       "l2": "working set per step (activations + weights, > 1 GB) exceeds the 126 MB L2; no flush",
                       "gflop_per_image_reference_count": gf},
            "clocks": m["clocks"],
            "e2e": {"value": m["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "ms_per_step": m["e2e_ms"],
                    "pipeline": "one upload per step; batch k+1 goes up on a copy stream behind step k" if m["use_graph"]
                    else "upload on the critical path"},
            "gpu_launches": m["launches"] * a.steps,
            "step_tflops_reference_count": step_tf,
            "step_frac_of_bf16_sustained": step_tf / pk["bf16_sustained"],
            "roofline": roof}
    single = rank == 0 and ws == 1
    if single and not a.no_cpu:
        try:
            cb = min(B, int(os.environ.get("EKL_CPU_BASELINE_BATCH", "8")))
            rate, kind, cores, sample, _, _ = cpu_step_rate(a.config, cb, 3, 1, float(os.environ.get("EKL_CPU_BUDGET_S", "45")),
                                                            min_steps=3)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed: %s" % ex}
        try:
            line["gpu_eager_baseline"] = gpu_eager_rate(a.config, B)
        except Exception as ex:  # noqa: BLE001
            line["gpu_eager_baseline"] = {"unit": UNIT, "failed": str(ex)[:200]}
    if single and not a.no_extra:
        # the other BASELINE configs at their own batch sizes: one short child run each (a fault there cannot take the
        # headline line down with it); the headline above is config `a.config`
        extra = {a.config: {"batch_per_gpu": B, "value": m["value"], "ms_per_step": m["ms_per_step"], "e2e": m["e2e_value"],
                            "frac_of_bf16_sustained": step_tf / pk["bf16_sustained"]}}
        m["graph"] = None
        del tr
        gc.collect()
        torch.cuda.empty_cache()
        ks, kw = min(a.steps, 10), min(max(a.warmup, 3), 3)
        for name in sorted(WORKLOAD):
            if name in extra:
                continue
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", name, "--steps", str(ks), "--warmup", str(kw),
                                    "--no-cpu", "--no-profile", "--no-extra"], capture_output=True, text=True, timeout=150)
                d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
                extra[name] = {"batch_per_gpu": d["config"]["batch_per_gpu"], "value": d["value"], "ms_per_step": d["ms_per_step"],
                               "e2e": d["e2e"]["value"], "frac_of_bf16_sustained": d["step_frac_of_bf16_sustained"]}
            except Exception as ex:  # noqa: BLE001
                extra[name] = {"failed": str(ex)[:200]}
        line["extra"] = {"all_configs": extra, "unit": UNIT,
                         "note": "each BASELINE config at its own batch per GPU, %d timed steps after %d warm-up, 1 GPU" % (ks, kw)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    teardown(ws, m.get("graph"))


# ------------------------------------------------------------------------------------------------ roofline
# kernel-name pattern -> family (first match wins).  Names are the demangled CUPTI kernel names.
FAMILIES = [
    ("conv_gemm_tc", r"conv_gemm_tc2?_kernel"),          # single-CTA and CTA-pair variants of the same tile program
    ("conv3x3_rw", r"conv3x3_rw_kernel"),
    ("conv_wgrad", r"conv_wgrad_"),
    ("splitk_finish", r"splitk_finish"),
    ("ekl_small", r"linear_bn_relu"),          # VC_NET's fused layers: not the feature-map BatchNorm passes accounted as "bn"
    ("batchnorm", r"bn_|col_stats"),
    ("adam", r"adam_step|adam_tick"),
    ("pack_weights", r"pack_"),
    ("capsule", r"caps_|dcaps_"),
    ("nccl", r"nccl"),
    ("ekl_small", r"ekl|anonymous namespace|<unnamed>"),
    ("library_gemm", r"gemm|cutlass|cublas|sm\d+_xmma|gemv"),
    ("torch_glue", r"."),
]
# ops-side accounting name -> family its flops / bytes belong to
ACCOUNT_FAMILY = {"conv_tc:generic": "conv_gemm_tc", "conv_tc:split": "conv_gemm_tc", "conv_tc:rw": "conv3x3_rw",
                  "conv_wgrad": "conv_wgrad", "bn": "batchnorm", "adam": "adam", "pack_weights": "pack_weights"}


def _family(name):
    for fam, pat in FAMILIES:
        if re.search(pat, name):
            return fam
    return "torch_glue"


def ncu_traffic(family):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed
    `ncu --set full` capture (profiles/r0N_conv_full.json, written by tools/ncu_summary.py) -> (bytes per launch,
    provenance); (None, None) when no capture is committed."""
    pdir = os.path.join(ROOT, "profiles")
    cands = sorted((f for f in os.listdir(pdir) if re.match(r"r\d+_conv_full\.json$", f)), reverse=True) if os.path.isdir(pdir) else []
    key = {"conv_gemm_tc": "conv_gemm_tc", "conv3x3_rw": "conv3x3_rw", "conv_wgrad": "conv_wgrad"}.get(family, family)
    for f in cands:
        rows = [r for r in json.load(open(os.path.join(pdir, f))) if key in r.get("kernel", "")]
        if rows:
            tot = sum(r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0) for r in rows)
            return tot / len(rows), {"launches_captured": len(rows), "source": "profiles/" + f}
    return None, None


def kernel_roofline(tr, gs, batch, pk, replays=3):
    """Per-kernel-family device time of the step AS BENCHED (CUDA-graph replay; CUPTI kernel records through
    torch.profiler, taken after the timed regions -- no bench value is measured under the profiler) joined with the
    algorithmic work of the same launches (one eager step with ops.ACCOUNT on: reference-count flops, executed flops and
    algorithmic bytes per library call, keyed by the kernel the call routes to).  `roofline` describes the dominant
    tensor-core family."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    from text2img_ekl_b200 import ops
    ops.ACCOUNT = []
    tr.train_step(batch)
    torch.cuda.synchronize()
    acc, ops.ACCOUNT = ops.ACCOUNT, None
    work = {}
    for name, ref_flops, exe_flops, nbytes in acc:
        d = work.setdefault(ACCOUNT_FAMILY.get(name, name), dict(ref=0.0, exe=0.0, bytes=0.0, calls=0))
        d["ref"] += ref_flops; d["exe"] += exe_flops; d["bytes"] += nbytes; d["calls"] += 1
    run = gs.replay if gs is not None else (lambda: tr.train_step(batch))
    run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(replays):
            run()
        torch.cuda.synchronize()
    fam = {}
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0.0)
        if t <= 0:
            continue
        d = fam.setdefault(_family(ev.key), dict(us=0.0, launches=0.0))
        d["us"] += t / replays
        d["launches"] += ev.count / replays
    total = sum(d["us"] for d in fam.values()) or 1.0
    out = {}
    for k, d in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        w = work.get(k, {})
        sec = d["us"] * 1e-6
        out[k] = {"us_per_step": round(d["us"], 1), "share": round(d["us"] / total, 4), "launches_per_step": round(d["launches"], 1),
                  "tflops_reference_count": round(w["ref"] / sec / 1e12, 1) if w.get("ref") else None,
                  "tflops_executed": round(w["exe"] / sec / 1e12, 1) if w.get("exe") else None,
                  "gbs_algorithmic": round(w["bytes"] / sec / 1e9, 1) if w.get("bytes") and not w.get("ref") else None,
                  "calls_accounted": w.get("calls")}
    tensor = [k for k in fam if work.get(k, {}).get("ref")]
    if not tensor:
        return {"families": out, "kernel_us_per_step": round(total, 1)}
    top = max(tensor, key=lambda k: fam[k]["us"])
    d, w = fam[top], work[top]
    ach = w["ref"] / (d["us"] * 1e-6) / 1e12
    conv_us = sum(fam[k]["us"] for k in tensor)
    conv_ref = sum(work[k]["ref"] for k in tensor)
    conv_exe = sum(work[k]["exe"] for k in tensor)
    traffic, traffic_src = ncu_traffic(top)
    return {"bound": "tensor", "kernel": top + "_kernel", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
            "frac": ach / pk["bf16_sustained"], "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": pk["src"] + " (sustained: kernel timed inside a long step)",
            "method": "CUPTI kernel durations of %d CUDA-graph replays of the benched step (after the timed regions) / "
                      "algorithmic flops of the same launches from one accounted eager step" % replays,
            "flop_counting": "reference dense-conv count (upsampled grid for up-convs, tiled code channels included)",
            "achieved_executed": w["exe"] / (d["us"] * 1e-6) / 1e12,
            "avg_launch_us": d["us"] / max(d["launches"], 1), "launches_per_step": d["launches"],
            # every accounted library call of this family is exactly one launch of it: a mismatch means the flops and the
            # device time describe different sets of launches (a kernel variant the family pattern misses)
            "calls_accounted": w["calls"], "launches_match_calls": abs(d["launches"] - w["calls"]) < 0.5,
            "all_conv_kernels": {"us_per_step": round(conv_us, 1), "tflops_reference_count": round(conv_ref / conv_us / 1e6, 1),
                                 "tflops_executed": round(conv_exe / conv_us / 1e6, 1),
                                 "frac_reference_count": conv_ref / conv_us / 1e6 / pk["bf16_sustained"]},
            "kernel_us_per_step": round(total, 1), "families": out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="3stages", choices=sorted(WORKLOAD))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and gpu_eager_baseline legs")
    ap.add_argument("--no-profile", action="store_true", help="skip the roofline pass")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other four BASELINE configs (`extra`, N=1 only)")
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
