#!/usr/bin/env python
"""Benchmark of the hot path: train samples/sec (forward + loss + backward + AdamW) at batch 4096 per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port)

Prints ONE JSON line (rank 0).  Keys follow the driver contract: metric/value/unit, ms_per_step, e2e (host buffers,
H2D + D2H inside the timed region), roofline (dominant kernel, live CUDA-event timing), cpu_baseline, clocks,
gpu_launches.  Workload = BASELINE.json configs[1]: RNA2DNAVAE train step at batch 4096 (synthetic 782 / 572 / 24).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "vae-los-angeles_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

DIMS = dict(A=782, B=572, S=24, L=20, E=32)
# algorithmic work per sample (fwd + bwd FLOP; compulsory HBM bytes) and parameter counts
WORK = {
    "rna2dna": dict(flop=3_013_120, bytes=7_872, params=538_124),
    "dna2rna": dict(flop=2_642_944, bytes=8_712, params=542_174),
    "multimodal": dict(flop=5_665_280, bytes=11_096, params=1_081_114),
    # directional autoencoders: the VAE figures minus one head per encoder (fwd + dgrad + wgrad = 6 FLOP per
    # head weight: 20 x (128 + 32) resp. 20 x (256 + 32) weights) and 80 B less output (latent only, no logvar)
    "rna2dna_ae": dict(flop=3_013_120 - 6 * 20 * (128 + 32), bytes=7_872 - 80, params=538_124 - 20 * (128 + 32) - 40),
    "dna2rna_ae": dict(flop=2_642_944 - 6 * 20 * (256 + 32), bytes=8_712 - 80, params=542_174 - 20 * (256 + 32) - 40),
}
METRIC = "train samples/sec (fwd+bwd+Adam) at batch 4096"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]), source="measured (MEASURED_PEAKS.json, burst)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The timed region is short (tens of ms), so the sampler
    polls NVML directly every ~2 ms from a thread (nvidia_ml_py); `nvidia-smi -lms` is the fallback."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = False
        self.proc, self.lines = None, []
        self.mode = None
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    idx = gpu_index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                              "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.mode = "smi"
                self.thread = threading.Thread(target=self._read, daemon=True)
                self.thread.start()
            except Exception:
                self.mode = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.mode is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["clock sampling unavailable"], samples=0)
        if self.mode == "nvml":
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = sorted(self.samples)
            return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                        samples=len(sm), source="nvml, 2 ms period")
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nme, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                    source="nvidia-smi -lms 20")


class NumaLocal:
    """Context manager: run the enclosed host allocations on the CPUs next to GPU `gpu_index` (NVML CPU affinity), so that the
    pinned staging buffers are first-touched on the GPU's own NUMA node; the previous affinity is restored on exit."""

    def __init__(self, gpu_index):
        self.gpu_index, self.prev, self.note = gpu_index, None, "not bound"

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                idx = int(vis.split(",")[self.gpu_index])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if cpus:
                self.prev = allowed
                os.sched_setaffinity(0, cpus)
                self.note = f"pinned buffers allocated on the GPU-local CPUs ({len(cpus)} of {len(allowed)})"
        except Exception as e:                      # no NVML / no affinity support: keep the default placement
            self.note = f"not bound ({type(e).__name__})"
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            os.sched_setaffinity(0, self.prev)
        return False


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_arm(workload, batch, steps, warmup, budget_s=25.0):
    import numpy as np
    from oracle import vae_oracle as vo
    # torchrun exports OMP_NUM_THREADS=1: set the BLAS pool explicitly to every core this process may use and report what
    # the pool really runs with (round 1's N > 1 lines silently timed one thread).
    want = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        threads = 1
    state = vo.init_state(workload, DIMS, seed=0)
    opt, step = vo.adamw_init(state)
    tpm, beta, site = vo.synthetic_batch(batch, DIMS, seed=0)
    data = dict(a=tpm, b=beta, site=site)
    eps, masks = vo.synthetic_noise(batch, DIMS, workload, seed=0)
    done, t_total = 0, 0.0
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, _, step = vo.train_step(workload, DIMS, state, opt, step, data, eps, masks)
        dt = time.perf_counter() - t0
        if i >= warmup:
            done += 1
            t_total += dt
            if t_total > budget_s:
                break
    sps = done * batch / t_total
    sample = f"{done} steps of batch {batch} after {warmup} warm-up steps, fp32 numpy (oracle/vae_oracle.py train_step)"
    return sps, threads, sample, 1e3 * t_total / done


def reference_main(args, rank):
    if rank != 0:
        return
    sps, threads, sample, ms = cpu_arm(args.workload, args.batch, min(args.steps, 20), min(args.warmup, 3))
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload} train step (fwd+loss+bwd+AdamW), batch {args.batch}, 782/572/24/latent 20",
                   "timing": "host wall clock (perf_counter)"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def shutdown_distributed(trainers, dev):
    """Release graphs that contain NCCL kernels, then tear the process group down; never hang at exit."""
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    threading.Timer(20.0, lambda: os._exit(0)).start()       # watchdog: the result line is already printed
    for t in trainers:
        t.close()
    torch.cuda.synchronize(dev)
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def train_throughput(workload, B, dev, steps, warmup):
    """Fused train step of another model kind at the same per-GPU batch (device-timed, CUDA graph replay)."""
    import torch
    from src.models import DNA2RNAVAE, MultiModalVAE, RNA2DNAVAE
    from src.models.directional_ae import DNA2RNAAE, RNA2DNAAE
    from vla_b200 import DeviceDataset, Trainer
    cls = {"rna2dna": RNA2DNAVAE, "dna2rna": DNA2RNAVAE, "multimodal": MultiModalVAE, "rna2dna_ae": RNA2DNAAE,
           "dna2rna_ae": DNA2RNAAE}[workload]
    torch.manual_seed(0)
    model = cls(DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"]).to(dev).train()
    ds = DeviceDataset.synthetic(B * 32, DIMS["A"], DIMS["B"], DIMS["S"], dev, seed=3)
    tr = Trainer(model, ds, B)
    for _ in range(max(warmup, 3)):
        tr.step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    w = WORK[workload]
    peaks = load_peaks()
    t_roof = max(w["flop"] * B / (peaks["bf16_tflops"] * 1e12), (w["bytes"] * B + 32 * w["params"]) / (peaks["hbm_gbs"] * 1e9))
    out = {"workload": f"{workload} train step, batch {B}", "value": steps * B / (ms * 1e-3), "unit": "samples/s",
           "ms_per_step": ms / steps, "step_roofline_frac": t_roof / (ms * 1e-3 / steps), "losses": tr.losses()}
    tr.close()
    del tr, model, ds
    torch.cuda.empty_cache()
    return out


def torch_eager_throughput(workload, B, dev, steps, warmup, mode):
    """The competitor a user would actually run: the reference's algorithm in stock PyTorch on the SAME B200
    (oracle/torch_eager.py: the ATen calls the reference's nn.Modules make + torch.optim.AdamW, eager, batches resident on the
    device).  mode: "fp32" (the reference's defaults), "tf32" (torch.backends.cuda.matmul.allow_tf32) or "bf16" (autocast)."""
    import torch
    from oracle import torch_eager as te
    from oracle import vae_oracle as vo
    from vla_b200 import DeviceDataset
    state = vo.init_state(workload, DIMS, seed=0)
    ds = DeviceDataset.synthetic(B * 8, DIMS["A"], DIMS["B"], DIMS["S"], dev, seed=3)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = mode in ("tf32", "bf16")
    try:
        tr = te.EagerTrainer(workload, state, dev, autocast=torch.bfloat16 if mode == "bf16" else None)

        def one(i):
            lo = (i % 8) * B
            return tr.step(ds.tpm[lo:lo + B], ds.beta[lo:lo + B], ds.site[lo:lo + B])

        for i in range(max(warmup, 3)):
            one(i)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            last = one(i)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        out = {"workload": f"stock PyTorch eager ({mode}) {workload} train step, batch {B}, same GPU (oracle/torch_eager.py)",
               "value": steps * B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms / steps, "final_loss": float(last)}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    del tr, ds
    torch.cuda.empty_cache()
    return out


def population_specs(n_models, first=0):
    """Members `first .. first + n_models` of the population: tri-modal VAEs with hyper-parameters drawn from the ranges of
    optimize_hyperparameters.py:71-76 (latent 10..100, embed 16/32/64, lr, weight decay, beta, gamma); member i is the same
    model whatever the sharding."""
    import numpy as np
    import torch
    from src.models import MultiModalVAE
    out = []
    for i in range(first, first + n_models):
        torch.manual_seed(1000 + i)
        r = np.random.default_rng(1000 + i)
        L, E = int(r.integers(10, 101)), int(r.choice([16, 32, 64]))
        out.append(dict(model=MultiModalVAE(DIMS["A"], DIMS["B"], DIMS["S"], L, embed_dim=E), seed=i,
                        lr=float(10 ** r.uniform(-5, -2)), weight_decay=float(10 ** r.uniform(-6, -3)),
                        beta_start=float(10 ** r.uniform(-4, -2)), gamma=float(r.uniform(0.5, 5.0))))
    return out


def population_throughput(dev, B, n_models, steps, warmup, compare=True, first=0, barrier=None):
    """BASELINE configs[4] on one GPU: `n_models` independent tri-modal VAEs stepped in LOCK-STEP (vla_b200.Population,
    grouped=True: every launch of the train step issued once for all members -- grouped GEMMs), against (compare=True) the same
    models as one graph per model on its own stream, and one after the other."""
    import torch
    from vla_b200 import DeviceDataset, Population
    ds = DeviceDataset.synthetic(B * 8, DIMS["A"], DIMS["B"], DIMS["S"], dev, seed=9)

    def timed(pop, fn, k):
        for _ in range(warmup):
            fn()
        pop.synchronize()
        if barrier:
            barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        for mem in pop.members:
            if mem.stream != torch.cuda.current_stream(dev):
                torch.cuda.current_stream(dev).wait_stream(mem.stream)
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / k

    pop = Population(population_specs(n_models, first), ds, B, device=dev, grouped=True)
    ms_group = timed(pop, lambda: pop.step(1), steps)
    losses = pop.losses()
    pop.close()
    del pop
    torch.cuda.empty_cache()
    ok = all(all(x == x and abs(x) < 1e30 for x in l) for l in losses)
    out = {"workload": f"population: {n_models} independent tri-modal VAEs (latent 10..100, embed 16/32/64), batch {B} each, one GPU, "
                       "lock-step grouped launches (vla_train_step_group)",
           "value": B * n_models / (ms_group * 1e-3), "unit": "samples/s (all members)", "ms_per_round": ms_group,
           "us_per_model_step": 1e3 * ms_group / n_models, "losses_finite": ok}
    if compare:
        pop = Population(population_specs(n_models, first), ds, B, device=dev, grouped=False)
        k = max(3, steps // 3)
        ms_streams = timed(pop, lambda: pop.step(1), k)

        def one_by_one():
            for mem in pop.members:                   # same graphs, serialised: every member's step waits for the previous member's
                with torch.cuda.stream(mem.stream):
                    mem.trainer.step()
                torch.cuda.current_stream(dev).wait_stream(mem.stream)
                for other in pop.members:
                    other.stream.wait_stream(torch.cuda.current_stream(dev))

        ms_seq = timed(pop, one_by_one, k)
        pop.close()
        del pop
        torch.cuda.empty_cache()
        out.update({"one_graph_per_model_on_streams_samples_per_s": B * n_models / (ms_streams * 1e-3),
                    "one_after_the_other_samples_per_s": B * n_models / (ms_seq * 1e-3),
                    "gain_over_one_after_the_other": ms_seq / ms_group, "gain_over_streams": ms_streams / ms_group})
    return out


def metrics_throughput(dev, rows, dim, steps, warmup):
    """Evaluation metrics of BASELINE configs[3]'s reconstructions (compare_directional_imputation.py:167-210) as one
    streaming kernel: 8 B of input per element, HBM-bound."""
    import ctypes as C
    import torch
    from vla_b200 import _lib
    from vla_b200.core import _ptr, _stream
    g = torch.Generator(device=dev); g.manual_seed(3)
    t = torch.rand(rows, dim, device=dev, generator=g)
    p = t + 0.1 * torch.randn(rows, dim, device=dev, generator=g)
    L = _lib.lib()
    out = torch.empty(8, dtype=torch.float64, device=dev)
    ws = torch.zeros(L.vla_metrics_workspace_bytes(rows), dtype=torch.uint8, device=dev)
    cos, pear = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    args = _lib.MetricsArgs(y_true=_ptr(t), y_pred=_ptr(p), rows=rows, dim=dim, cosine=_ptr(cos), pearson=_ptr(pear), out=_ptr(out),
                            workspace=_ptr(ws))
    for _ in range(warmup):
        _lib.check(L.vla_recon_metrics(C.byref(args), _stream()), "vla_recon_metrics")
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        _lib.check(L.vla_recon_metrics(C.byref(args), _stream()), "vla_recon_metrics")
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    peaks = load_peaks()
    gbs = 8.0 * rows * dim / (ms * 1e-3) / 1e9
    res = {"workload": f"reconstruction metrics (MAE/MSE/R2/cosine/Pearson), {rows} x {dim} fp32, one kernel", "value": rows / (ms * 1e-3),
           "unit": "samples/s", "ms_per_step": ms, "bound": "hbm", "achieved_gb_per_s": gbs, "peak_gb_per_s": peaks["hbm_gbs"],
           "roofline_frac": gbs / peaks["hbm_gbs"], "l2": f"inputs {8.0 * rows * dim / 1e6:.0f} MB > L2", "mse": out.tolist()[1]}
    del t, p
    torch.cuda.empty_cache()
    return res


def inference_throughput(dev, batch, steps, warmup):
    """BASELINE configs[3]: tri-modal cross-modal inference `model(a=x)` in eval mode (BatchNorm running statistics, no
    dropout, epsilon still sampled), all three decoders, fp32 outputs written; vla_forward replayed from a CUDA graph."""
    import ctypes as C
    import torch
    from src.models import MultiModalVAE
    from vla_b200 import _lib
    from vla_b200.core import _ptr, _stream
    torch.manual_seed(0)
    model = MultiModalVAE(DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"]).to(dev).eval()
    core = model._ensure_core()
    g = torch.Generator(device=dev); g.manual_seed(5)
    x = torch.log1p(-20.0 * torch.log1p(-torch.rand(batch, DIMS["A"], device=dev, generator=g)))
    outs = [torch.empty(batch, n, device=dev) for n in (DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"], DIMS["L"])]
    L = _lib.lib()

    def call(refresh):
        args = _lib.ForwardArgs(params=_ptr(core.arena), buffers=_ptr(core.buffers), counters=_ptr(core.counters), x_a=_ptr(x),
                                x_b=None, site=None, batch=batch, train=0, refresh_shadows=refresh, eps=None, keep_masks=None,
                                seed=1, offset=0, recon_a=_ptr(outs[0]), recon_b=_ptr(outs[1]), recon_c=_ptr(outs[2]),
                                mu=_ptr(outs[3]), logvar=_ptr(outs[4]))
        _lib.check(L.vla_forward(core.handle, C.byref(args), _stream()), "vla_forward")

    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        call(1)
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        call(0)
    for _ in range(warmup):
        graph.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    peaks = load_peaks()
    # a-only eval forward = 639 744 MACs, 3 128 B in + 5 672 B out per sample
    t_roof = max(2 * 639_744 * batch / (peaks["bf16_tflops"] * 1e12), 8_800 * batch / (peaks["hbm_gbs"] * 1e9))
    ok = bool(torch.isfinite(outs[1]).all().item()) and float(outs[1].min()) >= 0.0 and float(outs[1].max()) <= 1.0
    out = {"workload": f"multimodal eval inference model(a=x), batch {batch}", "value": steps * batch / (ms * 1e-3), "unit": "samples/s",
           "ms_per_step": ms / steps, "step_roofline_frac": t_roof / (ms * 1e-3 / steps), "bound": "hbm", "outputs_valid": ok}
    del graph, outs, x, model
    torch.cuda.empty_cache()
    return out


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from src.models import DNA2RNAVAE, MultiModalVAE, RNA2DNAVAE
    from vla_b200 import DeviceDataset, Trainer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1 or os.environ.get("VLA_FORCE_DP") == "1":    # (the override measures the exchange's local cost at world 1)
        os.environ.setdefault("MASTER_PORT", "29533"); os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    from src.models.directional_ae import DNA2RNAAE, RNA2DNAAE
    cls = {"rna2dna": RNA2DNAVAE, "dna2rna": DNA2RNAVAE, "multimodal": MultiModalVAE, "rna2dna_ae": RNA2DNAAE,
           "dna2rna_ae": DNA2RNAAE}[args.workload]
    torch.manual_seed(0)                       # identical replicas on every rank
    model = cls(DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"]).to(dev).train()
    B = args.batch
    n_batches = args.resident_batches
    ds = DeviceDataset.synthetic(B * n_batches, DIMS["A"], DIMS["B"], DIMS["S"], dev, seed=1000 + rank)
    trainer = Trainer(model, ds, B, lr=5e-4, weight_decay=1e-5, beta_kl=1e-3, gamma=1.0, seed=rank, process_group=pg,
                      exchange=args.exchange)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        trainer.step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        trainer.step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    losses = trainer.losses()
    dp_trace = trainer.exchange_trace()
    assert all(x == x and abs(x) < 1e30 for x in losses), f"non-finite loss {losses}"
    value = args.steps * B * world / (ms * 1e-3)

    # ---- end to end: pinned host buffers -> H2D -> step -> D2H of the loss, double buffered -------------------
    e2e = None
    host = [DeviceDataset.synthetic(B, DIMS["A"], DIMS["B"], DIMS["S"], dev, seed=77 + i + 10 * rank) for i in range(4)]
    with NumaLocal(local_rank) as numa:
        pinned = [(h.tpm.cpu().pin_memory(), h.beta.cpu().pin_memory(), h.site.cpu().pin_memory()) for h in host]
    del host
    slots = [DeviceDataset.synthetic(B, DIMS["A"], DIMS["B"], DIMS["S"], dev, seed=5 + i) for i in range(2)]
    tr2 = Trainer(model, slots, B, lr=5e-4, weight_decay=1e-5, beta_kl=1e-3, gamma=1.0, seed=rank, process_group=pg,
                  exchange=args.exchange)
    with NumaLocal(local_rank):
        loss_host = [torch.zeros(4).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    h2d_bytes = sum(x.numel() * x.element_size() for x in pinned[0])
    k_e2e = max(8, min(args.steps, 64))

    def e2e_loop(k, timed):
        copied = [torch.cuda.Event() for _ in range(k)]
        done = [torch.cuda.Event() for _ in range(k)]
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream):
            s0.record()
        for i in range(k):
            slot = slots[i % 2]
            src = pinned[i % len(pinned)]
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(done[i - 2])          # the slot's previous step has consumed it
                slot.tpm.copy_(src[0], non_blocking=True)
                slot.beta.copy_(src[1], non_blocking=True)
                slot.site.copy_(src[2], non_blocking=True)
                copied[i].record()
            main_stream.wait_event(copied[i])
            tr2.step(i % 2)
            loss_host[i % 2].copy_(tr2.loss_out, non_blocking=True)
            done[i].record()
        s1.record()
        torch.cuda.synchronize(dev)
        return s0.elapsed_time(s1)

    e2e_loop(4, False)          # warm-up: captures both graphs
    barrier()
    ms2 = e2e_loop(k_e2e, True)
    t = torch.tensor([ms2], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms2 = float(t.item())
    e2e = {"value": k_e2e * B * world / (ms2 * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes * world,
           "d2h_bytes_per_step": 16 * world, "steps": k_e2e, "ms_per_step": ms2 / k_e2e,
           "h2d_gb_per_s": h2d_bytes / (ms2 / k_e2e * 1e-3) / 1e9, "host_numa": numa.note,
           "how": "pinned host batch -> cudaMemcpyAsync H2D (copy stream, double buffered) -> fused step graph -> 16 B loss D2H; "
                  "bound by the host->device link (h2d_gb_per_s), the step itself takes ms_per_step of the device-resident line"}

    # ---- per-launch timing (eager, CUDA events inside the library) -> dominant kernel roofline ----------------
    peaks = load_peaks()
    prof_steps = 10
    entries = trainer.profile(prof_steps)
    agg = {}
    for name, pms, fl, by in entries:
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += pms; a[2] += fl; a[3] += by
    # an event pair with nothing between it measures the timing overhead itself: subtract it from every launch
    empty = agg.pop("_empty_pair", None)
    overhead_ms = (empty[1] / empty[0]) if empty else 0.0
    for a in agg.values():
        a[1] = max(a[1] - overhead_ms * a[0], 0.05e-3 * a[0])
    total_ms = sum(a[1] for a in agg.values()) or 1.0
    kernels = []
    for name, (cnt, pms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        avg_ms, fl, by = pms / cnt, fl / cnt, by / cnt
        t_t = fl / (peaks["bf16_tflops"] * 1e12)
        t_h = by / (peaks["hbm_gbs"] * 1e9)
        kernels.append(dict(name=name, launches_per_step=cnt / prof_steps, avg_us=1e3 * avg_ms, share=pms / total_ms,
                            flops=fl, bytes=by, bound="tensor" if t_t > t_h else "hbm",
                            frac=max(t_t, t_h) / (avg_ms * 1e-3) if avg_ms > 0 else None))
    # the decoder weight gradients run on the low-priority side branch beside the encoder backward (64 CTAs, off the critical
    # path: their event-pair time is stretched by design); the dominant kernel is the longest launch of the main chain
    for k in kernels:
        k["side_branch"] = world == 1 and k["name"] in ("wgrad_dec", "adamw_dec")
    top = next(k for k in kernels if not k["side_branch"])
    w = WORK[args.workload]
    if top["bound"] == "tensor":
        ach, peak, unit = top["flops"] / (top["avg_us"] * 1e-6) / 1e12, peaks["bf16_tflops"], "TFLOP/s"
    else:
        ach, peak, unit = top["bytes"] / (top["avg_us"] * 1e-6) / 1e9, peaks["hbm_gbs"], "GB/s"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if os.path.exists(tpath) and args.workload == "rna2dna" and B == 4096:
        with open(tpath) as f:
            traffic = json.load(f).get(top["name"])          # DRAM bytes per launch from the committed ncu --set full capture
    roofline = {"kernel": top["name"], "bound": top["bound"], "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "traffic": traffic, "peak_source": peaks["source"], "avg_us": top["avg_us"], "share_of_step": top["share"],
                "event_pair_overhead_us": 1e3 * overhead_ms,
                "how": "CUDA event pair around each launch, recorded as nodes of the captured step graph, 10 replays; the duration of "
                       "an empty event pair is subtracted (vla_profile_*); achieved = algorithmic flops or bytes of the launch / that time"}
    phases = None
    if world == 1:
        # inside the chain launches: per-phase spans from the %globaltimer stamps every CTA writes (one eager step)
        try:
            phases = [{"chain": c["name"], "ctas": c["ctas"], "span_us": round(c["span_us"], 2),
                       "phases": [{k: (round(v, 2) if isinstance(v, float) else v) for k, v in ph.items()} for ph in c["phases"]]}
                      for c in trainer.timeline()]
        except Exception as e:                       # a diagnostic, never fatal to the measurement
            phases = {"error": f"{type(e).__name__}: {e}"}
    t_tensor = w["flop"] * B / (peaks["bf16_tflops"] * 1e12)
    t_hbm = (w["bytes"] * B + 32 * w["params"]) / (peaks["hbm_gbs"] * 1e9)
    step_s = ms * 1e-3 / args.steps
    step_roofline = {"bound": "tensor" if t_tensor > t_hbm else "hbm", "t_tensor_us": 1e6 * t_tensor, "t_hbm_us": 1e6 * t_hbm,
                     "t_step_us": 1e6 * step_s, "frac": max(t_tensor, t_hbm) / step_s,
                     "flop_per_sample": w["flop"], "bytes_per_sample": w["bytes"], "param_bytes_per_step": 32 * w["params"]}
    launches_per_step = sum(k["launches_per_step"] for k in kernels)

    # ---- the other BASELINE.json configurations, short runs (reported under "also"; not the headline) ----------------
    also = []
    extra_trainers = []
    if world == 1 and not args.no_also:
        for wl in ("multimodal", "dna2rna", "rna2dna", "rna2dna_ae", "dna2rna_ae"):
            if wl == args.workload:
                continue
            also.append(train_throughput(wl, B, dev, steps=60, warmup=5))
        for mode in ("fp32", "tf32", "bf16"):
            also.append(torch_eager_throughput(args.workload, B, dev, steps=40, warmup=5, mode=mode))
        also.append(inference_throughput(dev, batch=args.infer_batch, steps=10, warmup=3))
        also.append(metrics_throughput(dev, args.infer_batch, DIMS["A"], steps=10, warmup=3))
        # BASELINE configs[4]: 40 models per GPU (320 = 64 trials x 5 folds over 8 GPUs)
        also.append(population_throughput(dev, B, n_models=40, steps=12, warmup=3))
        also.append(population_throughput(dev, 32, n_models=40, steps=60, warmup=5))       # the reference's default batch size

    # ---- data parallel: driver-visible parity of the replicas + BASELINE configs[2] (dna2rna, global batch = world x B) ----
    dp_check = None
    if world > 1:
        arena = trainer.core.arena.detach()
        digest = torch.stack([arena.double().sum(), arena.double().abs().sum(), arena.double().pow(2).sum()])
        lo, hi = digest.clone(), digest.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        # what the exchange delivers vs the per-shard losses: a fresh replica pair per rank (same initial state everywhere),
        # one step with the exchange and one without on the same shard batch, same Philox streams; the delivered loss must
        # be the sum over the ranks of the solo losses
        torch.manual_seed(123)
        m_dp = cls(DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"]).to(dev).train()
        m_solo = cls(DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"]).to(dev).train()
        m_solo.load_state_dict(m_dp.state_dict())
        ds_c = DeviceDataset.synthetic(B * 2, DIMS["A"], DIMS["B"], DIMS["S"], dev, seed=500 + rank)
        t_dp = Trainer(m_dp, ds_c, B, lr=5e-4, weight_decay=1e-5, beta_kl=1e-3, gamma=1.0, seed=rank, process_group=pg,
                       exchange=args.exchange, use_graph=False)
        t_solo = Trainer(m_solo, ds_c, B, lr=5e-4, weight_decay=1e-5, beta_kl=1e-3, gamma=1.0, seed=rank, use_graph=False)
        t_dp.step()
        t_solo.step()
        barrier()
        own = torch.tensor(list(t_solo.losses()), dtype=torch.float64, device=dev)
        dist.all_reduce(own, op=dist.ReduceOp.SUM)
        delivered = torch.tensor(list(t_dp.losses()), dtype=torch.float64, device=dev)
        t_solo.close()
        extra_trainers.append(t_dp)
        dp_check = {"replicas_identical": bool(torch.equal(lo, hi)),
                    "param_digest": [float(x) for x in digest.tolist()],
                    "loss_sum_over_shards": [float(x) for x in own.tolist()],
                    "loss_delivered_by_exchange": [float(x) for x in delivered.tolist()],
                    "loss_rel_diff": float(((own - delivered).abs() / own.abs().clamp_min(1e-6)).max())}
        if not args.no_also:
            torch.manual_seed(0)
            m2 = DNA2RNAVAE(DIMS["A"], DIMS["B"], DIMS["S"], DIMS["L"]).to(dev).train()
            ds2 = DeviceDataset.synthetic(B * 16, DIMS["A"], DIMS["B"], DIMS["S"], dev, seed=2000 + rank)
            t2 = Trainer(m2, ds2, B, lr=5e-4, weight_decay=1e-5, beta_kl=1e-3, gamma=1.0, seed=rank, process_group=pg, exchange=args.exchange)
            for _ in range(5):
                t2.step()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(100):
                t2.step()
            a1.record()
            barrier()
            tt = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            also.append({"workload": f"dna2rna train step (BASELINE configs[2]), data parallel over {world} GPUs, global batch {B * world}",
                         "value": 100 * B * world / (float(tt.item()) * 1e-3), "unit": "samples/s", "ms_per_step": float(tt.item()) / 100,
                         "losses": t2.losses()})
            extra_trainers.append(t2)
            # BASELINE configs[4]: the population sharded over the GPUs (members rank * 40 .. rank * 40 + 39 here; replicas
            # only, no data-path collective), every shard in lock-step grouped launches
            per_gpu = 40
            pl = population_throughput(dev, 32, n_models=per_gpu, steps=40, warmup=5, compare=False, first=rank * per_gpu, barrier=barrier)
            tp = torch.tensor([pl["ms_per_round"]], device=dev, dtype=torch.float64)
            okf = torch.tensor([1.0 if pl["losses_finite"] else 0.0], device=dev, dtype=torch.float64)
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            also.append({"workload": f"population (BASELINE configs[4]): {per_gpu * world} independent tri-modal VAEs, {per_gpu} per GPU over {world} GPUs "
                                     "(sharded by index, no collective), batch 32 each, lock-step grouped launches",
                         "value": 32 * per_gpu * world / (float(tp.item()) * 1e-3), "unit": "samples/s (all members, all GPUs)",
                         "ms_per_round_max_over_ranks": float(tp.item()), "losses_finite": bool(okf.item() > 0)})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sps, threads, sample, _ = cpu_arm(args.workload, B, 10, 2, budget_s=20.0)
        cpu = {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload} train step (fwd+loss+bwd+AdamW), batch {B} per GPU, 782/572/24/latent 20",
                       "global_batch": B * world, "parallelism": (f"dp{world}, batch-sharded replicas, one all-reduce(SUM) of the gradient arena per step: " +
                                       ("own kernel over NVLink peer memory (csrc/dp_exchange.cu)" if args.exchange == "p2p"
                                        else "NCCL all_reduce")) if world > 1 else "single",
                       "l2": f"inputs larger than L2: {n_batches} resident batches ({B * n_batches} rows, "
                             f"{B * n_batches * (DIMS['A'] + DIMS['B']) * 4 / 1e6:.0f} MB) visited in turn",
                       "arithmetic": "bf16 tensor-core operands, fp32 accumulate, fp32 master weights / loss / AdamW",
                       "graph": "CUDA graph replay per step, no host sync in the timed region"},
            "e2e": e2e, "roofline": roofline, "step_roofline": step_roofline, "kernels": kernels[:8], "timeline": phases,
            "gpu_launches": int(round(launches_per_step * args.steps)), "launches_per_step": launches_per_step,
            "cpu_baseline": cpu, "dp_exchange_trace": dp_trace, "dp_check": dp_check, "clocks": clocks, "final_losses": losses, "also": also,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_distributed([trainer, tr2] + extra_trainers, dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rna2dna", choices=sorted(WORK))
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--resident-batches", type=int, default=64)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="data-parallel gradient exchange (N > 1)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads (tri-modal, dna2rna, inference)")
    ap.add_argument("--infer-batch", type=int, default=262144)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # stdout carries ONE line (the JSON result).  NCCL prints its version banner with printf on file descriptor 1 whatever
    # NCCL_DEBUG says on some boxes: point descriptor 1 at stderr for native code and keep the real stdout for Python's print.
    os.environ["NCCL_DEBUG"] = os.environ.get("VLA_NCCL_DEBUG", "WARN")
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_out, "w", buffering=1)
    if args.impl == "reference":
        reference_main(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
