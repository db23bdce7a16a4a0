#!/usr/bin/env python
"""Benchmark of the structured-layer hot path (BASELINE.json metric: structured-layer fwd+bwd samples/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sss|lr|psm|hmat|ldr] [--impl ours|reference]

A "step" is one forward + backward (parameter gradients only, as training_helpers.py:141-144 produces)
of the layer over one batch of synthetic features that are already resident in HBM; for N > 1 the
batch is sharded over the ranks (data parallel, strong scaling: the global batch is fixed) and the step
ends with one NCCL all-reduce of the flat gradient buffer.  Prints ONE JSON line (rank 0).

`--impl reference` times the reference's CPU path (oracle port: the same torch CPU ops the reference
layer issues, autograd backward) on the box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np
import torch

HBM_FALLBACK_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback
BF16_FALLBACK_TFLOPS = 1590.0
FP32_FMA_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.4: 148 SMs x 128 FMA lanes x 2 flop x max clock


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_tflops=float(p["bf16_tflops"]),
                    bf16_tflops_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=HBM_FALLBACK_GBS, bf16_tflops=BF16_FALLBACK_TFLOPS, bf16_tflops_sustained=1400.0, source="fallback")


# ----------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------
class SSSWorkload:
    """BASELINE config C5 (= C1's layer): SSSLayer 4096 -> 1000, 500 stages, statespace dim 16, fp32,
    global batch 65536 sharded over the ranks."""
    name = "sss"
    dtype = "f32"
    input_dim, output_dim, nb_states, statespace_dim = 4096, 1000, 500, 16
    global_batch = 65536
    cpu_sample_batch = 256
    bytes_per_sample_kernel = 4 * (4096 + 1000)          # one kernel (fwd: read x, write y; bwd: read gy, re-read x)
    bytes_per_sample = 2 * bytes_per_sample_kernel        # 40 768 B / sample fwd+bwd
    flop_per_sample = 2277504
    bound = "hbm"
    # dram__bytes_read.sum + dram__bytes_write.sum per launch at local batch 65 536, from the `ncu --set full` capture
    # profiles/r1q_sss_tc_top_kernels.md (cold-cache replays of the same bench command, final kernels of round 1)
    traffic_bytes = {"sss_tc_local_gemm_kernel": 1.0775e9 + 0.5094e9, "sss_tc_chain_fwd_kernel": 0.7607e9 + 0.7648e9,
                     "sss_tc_chain_bwd_kernel": 0.5354e9 + 0.2493e9, "sss_tc_grad_gemm_kernel": 1.9521e9 + 0.0074e9}

    def describe(self):
        return dict(workload="C5: SSS 4096->1000, 500 stages, statespace 16, fp32 fwd+bwd (param grads), global batch 65536",
                    global_batch=self.global_batch, timed_inputs="resident in HBM; 1 GiB feature matrix per step >> 126 MB L2")

    def system(self):
        from structurednets_b200.synth import random_mixed_system
        return random_mixed_system(self.input_dim, self.output_dim, self.nb_states, self.statespace_dim, seed=5000)

    def make_layer(self, device):
        from structurednets_b200.layers.sss_layer import SSSLayer
        layer = SSSLayer(self.input_dim, self.output_dim, 0.105, nb_states=self.nb_states, initial_system_approx=self.system())
        assert layer.statespace_dim == self.statespace_dim
        return layer.to(device)

    def make_inputs(self, batch, device, seed):
        g = torch.Generator(device=device).manual_seed(seed)
        x = torch.rand((batch, self.input_dim), device=device, generator=g) * 2 - 1
        gy = (torch.rand((batch, self.output_dim), device=device, generator=g) * 2 - 1) / batch
        labels = torch.randint(0, self.output_dim, (batch,), device=device, generator=g)
        return x, gy, labels

    def cpu_step_fn(self, batch):
        """The reference's CPU path for this workload (oracle port of sss_layer.py:99-131 + autograd)."""
        from oracle import layers_cpu as O
        sysm = self.system()
        f32 = lambda m: torch.tensor(np.asarray(m)).float().requires_grad_(True)
        cs, acs = sysm.causal_system.stages, sysm.anticausal_system.stages
        lists = [[f32(s.A_matrix) for s in cs], [f32(s.B_matrix) for s in cs], [f32(s.C_matrix) for s in cs],
                 [f32(s.D_matrix) for s in cs], [f32(s.A_matrix) for s in acs], [f32(s.B_matrix) for s in acs],
                 [f32(s.C_matrix) for s in acs]]
        bias = torch.zeros(self.output_dim, requires_grad=True)
        rng = np.random.default_rng(5000)
        x = torch.tensor(rng.uniform(-1, 1, size=(batch, self.input_dim)).astype(np.float32))
        gy = torch.tensor(rng.uniform(-1, 1, size=(batch, self.output_dim)).astype(np.float32)) / batch
        params = [p for l in lists for p in l] + [bias]

        def step():
            for p in params:
                p.grad = None
            y = O.sss_forward(x, *lists, bias, sysm.dims_in, sysm.dims_out)
            y.backward(gy)
        return step


class _DenseInputsMixin:
    def make_inputs(self, batch, device, seed):
        g = torch.Generator(device=device).manual_seed(seed)
        x = torch.rand((batch, self.input_dim), device=device, generator=g) * 2 - 1
        gy = (torch.rand((batch, self.output_dim), device=device, generator=g) * 2 - 1) / batch
        labels = torch.randint(0, self.output_dim, (batch,), device=device, generator=g)
        return x, gy, labels

    def _cpu_xy(self, batch, seed):
        rng = np.random.default_rng(seed)
        x = torch.tensor(rng.uniform(-1, 1, size=(batch, self.input_dim)).astype(np.float32))
        gy = torch.tensor(rng.uniform(-1, 1, size=(batch, self.output_dim)).astype(np.float32)) / batch
        return x, gy


class LRWorkload(_DenseInputsMixin):
    """BASELINE config C2 shape: LRLayer 2048 -> 1000, rank 128, batch 8192 (fp32 CUDA-core path of this round)."""
    name, dtype, bound = "lr", "f32", "hbm"
    input_dim, output_dim, rank = 2048, 1000, 128
    global_batch, cpu_sample_batch = 8192, 8192
    bytes_per_sample_kernel = 4 * (2048 + 1000)
    bytes_per_sample = 2 * bytes_per_sample_kernel
    flop_per_sample = 1816576
    kernels = ("gemm_f32_kernel", "gemm_f32_kernel")

    def describe(self):
        return dict(workload="C2 shape: low-rank 2048->1000 rank 128, fp32 fwd+bwd, batch 8192 (bf16 tcgen05 path: next)",
                    global_batch=self.global_batch, timed_inputs="resident in HBM")

    def make_layer(self, device):
        from structurednets_b200.layers.lr_layer import LRLayer
        np.random.seed(2000)
        layer = LRLayer(self.input_dim, self.output_dim, 0.1906)
        assert layer.left_lr.shape[1] == self.rank
        return layer.to(device)

    def cpu_step_fn(self, batch):
        from oracle import layers_cpu as O
        np.random.seed(2000)
        lim = lambda shape: np.sqrt(6 / sum(shape))
        L = torch.tensor(np.random.uniform(-lim((1000, 128)), lim((1000, 128)), (1000, 128)).astype(np.float32), requires_grad=True)
        R = torch.tensor(np.random.uniform(-lim((128, 2048)), lim((128, 2048)), (128, 2048)).astype(np.float32), requires_grad=True)
        b = torch.zeros(1000, requires_grad=True)
        x, gy = self._cpu_xy(batch, 2000)

        def step():
            for p in (L, R, b):
                p.grad = None
            O.lr_forward(x, L, R, b).backward(gy)
        return step


class LRBf16Workload(LRWorkload):
    """BASELINE config C2: LRLayer 2048 -> 1000, rank 128, bf16 fwd+bwd (tcgen05 / TMEM / TMA), fp32 master parameters."""
    name, dtype, bound = "lr_bf16", "bf16", "hbm"
    bytes_per_sample_kernel = 2 * (2048 + 1000)
    bytes_per_sample = 2 * bytes_per_sample_kernel        # 12 192 B / sample
    kernels = ("gemm_bf16_tc_kernel", "gemm_bf16_tc_kernel")

    def describe(self):
        return dict(workload="C2: low-rank 2048->1000 rank 128, bf16 fwd+bwd on tcgen05 (fp32 accumulate, fp32 master params), batch 8192",
                    global_batch=self.global_batch, timed_inputs="resident in HBM")

    def make_inputs(self, batch, device, seed):
        x, gy, labels = super().make_inputs(batch, device, seed)
        return x.bfloat16(), gy.bfloat16(), labels


class PSMWorkload(_DenseInputsMixin):
    """BASELINE config C3: PSMLayer 4096 -> 1000, 3 factors, nb_params_share 0.1 => 409 599 nnz, batch 8192, intended
    factor order."""
    name, dtype, bound = "psm", "f32", "hbm"
    input_dim, output_dim = 4096, 1000
    global_batch, cpu_sample_batch = 8192, 512
    bytes_per_sample_kernel = 4 * (4096 + 1000)
    bytes_per_sample = 2 * bytes_per_sample_kernel
    flop_per_sample = 2184528
    kernels = ("psm_fwd_kernel", "psm_bwd_kernel")

    def describe(self):
        return dict(workload="C3: PSM 4096->1000, 3 sparse factors, 409 599 nnz (share 0.1), fp32 fwd+bwd, batch 8192",
                    global_batch=self.global_batch, timed_inputs="resident in HBM")

    def factors(self):
        import scipy.sparse
        shapes = [(1000, 4096), (4096, 4096), (4096, 4096)]
        out = []
        for k, (r, c) in enumerate(shapes):
            m = scipy.sparse.random(r, c, density=136533 / (r * c), random_state=3000 + k, format="csr", dtype=np.float64)
            m.data = np.random.default_rng(3000 + k).uniform(-0.05, 0.05, size=m.data.shape)
            out.append(m)
        return out

    def make_layer(self, device):
        from structurednets_b200.layers.psm_layer import PSMLayer
        return PSMLayer(self.input_dim, self.output_dim, sparse_matrices=self.factors()).to(device)

    def cpu_step_fn(self, batch):
        from oracle import layers_cpu as O
        fs = [torch.tensor(f.toarray()).float().requires_grad_(True) for f in self.factors()]
        b = torch.zeros(1000, requires_grad=True)
        x, gy = self._cpu_xy(batch, 3000)

        def step():   # the reference's default path: dense mm chain (psm_layer.py:51-58), corrected factor order
            for p in fs + [b]:
                p.grad = None
            O.psm_forward(x, fs, b).backward(gy)
        return step


class HMatWorkload(_DenseInputsMixin):
    """BASELINE config C4-H: HMatLayer 2048 -> 1000, eta 0.5, min block 2 => 2560 leaves of rank min(6, dim-1), batch 8192."""
    name, dtype, bound = "hmat", "f32", "hbm"
    input_dim, output_dim = 2048, 1000
    global_batch, cpu_sample_batch = 8192, 64
    bytes_per_sample_kernel = 4 * (2048 + 1000)
    bytes_per_sample = 2 * bytes_per_sample_kernel
    flop_per_sample = 2130000
    kernels = ("hmat_fwd_kernel", "hmat_bwd_kernel")

    def describe(self):
        return dict(workload="C4-H: H-matrix 2048->1000, eta 0.5, 2560 leaves (rank <= 6), fp32 fwd+bwd, batch 8192",
                    global_batch=self.global_batch, timed_inputs="resident in HBM")

    def hmatrix(self):
        from structurednets_b200.hmatrix import HMatrix, build_hmat_block_cluster_tree
        rng = np.random.default_rng(4000)
        tree = build_hmat_block_cluster_tree((1000, 2048), eta=0.5, min_block_size=2)
        for leaf in tree.get_all_leaf_elements():
            rows, cols = len(leaf.row_range), len(leaf.col_range)
            k = min(6, min(rows, cols) - 1)
            leaf.set_hmatrix_component(torch.tensor(rng.uniform(-1, 1, (rows, k)) / np.sqrt(k)).float(),
                                       torch.tensor(rng.uniform(-1, 1, (k, cols)) / np.sqrt(cols)).float())
        return HMatrix(tree, shape=(1000, 2048))

    def make_layer(self, device):
        from structurednets_b200.layers.hmat_layer import HMatLayer
        return HMatLayer(self.input_dim, self.output_dim, 0.2, initial_hmatrix=self.hmatrix()).to(device)

    def cpu_step_fn(self, batch):
        from oracle import layers_cpu as O
        hm = self.hmatrix()
        comps = [(c.row_range.start, c.row_range.stop, c.col_range.start, c.col_range.stop, c.left_lr.detach().clone().requires_grad_(True),
                  c.right_lr.detach().clone().requires_grad_(True)) for c in hm.get_all_hmatrix_components() if c.are_low_rank_components_set()]
        b = torch.zeros(1000, requires_grad=True)
        x, gy = self._cpu_xy(batch, 4000)

        def step():
            for c in comps:
                c[4].grad = None; c[5].grad = None
            O.hmat_forward(x, comps, b, 1000).backward(gy)
        return step


class LDRWorkload(_DenseInputsMixin):
    """BASELINE config C4-L: LDRLayer 2048 -> 2048 (square only), share 0.1 => displacement rank 99, batch 8192."""
    name, dtype, bound = "ldr", "f64 series + f32 apply", "hbm"
    input_dim, output_dim = 2048, 2048
    global_batch, cpu_sample_batch = 8192, 0
    bytes_per_sample_kernel = 4 * (2048 + 2048)
    bytes_per_sample = 2 * bytes_per_sample_kernel
    flop_per_sample = 16.8e6
    kernels = ("gemm_f32_kernel", "gemm_f32_kernel")

    def describe(self):
        return dict(workload="C4-L: LDR 2048->2048 (reference is square-only), displacement rank 99, batch 8192; the reference's own "
                             "construction is O(r n^4 log n) and cannot run at this size", global_batch=self.global_batch,
                    timed_inputs="resident in HBM")

    def make_layer(self, device):
        from structurednets_b200.layers.ldr_layer import LDRLayer
        np.random.seed(4100)
        layer = LDRLayer(2048, 2048, 0.1)
        assert layer.representation_matrices[2].shape[1] == 99
        return layer.to(device)

    def cpu_step_fn(self, batch):
        return None


WORKLOADS = {"sss": SSSWorkload, "lr": LRWorkload, "lr_bf16": LRBf16Workload, "psm": PSMWorkload, "hmat": HMatWorkload, "ldr": LDRWorkload}



# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._thr = None

    def _run_nvml(self):
        """NVML (nvidia_ml_py) sampling every ~5 ms: the timed region of a fast workload is shorter than one nvidia-smi call."""
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.gpu_index)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        bits = (("hw_slowdown", N.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", N.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksThrottleReasonSwPowerCap))
        while not self._stop.is_set():
            sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            try:
                pw = N.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            self.rows.append([str(self.gpu_index), str(sm), str(mx), str(pw)] + ["Active" if (r & b) else "Not Active" for _, b in bits])
            self._stop.wait(0.005)

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def time_cpu_baseline(wl, steps, warmup, budget_s=25.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B = wl.cpu_sample_batch
    step = wl.cpu_step_fn(B)
    for _ in range(warmup):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    med = float(np.median(times))
    return dict(value=B / med, unit="samples/s", cores=cores, kind="port",
                sample="%d steps of batch %d (median %.1f ms/step) of the same layer on %s" % (len(times), B, med * 1e3, cpu_model())), med, len(times)


# ----------------------------------------------------------------------------------------------
def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not wl.cpu_sample_batch:
        print(json.dumps(dict(impl="reference", unavailable="the reference cannot run this size on a CPU (O(r n^4 log n) weight build)")))
        return
    base, med, n = time_cpu_baseline(wl, steps=max(args.steps, 2), warmup=max(args.warmup, 1), budget_s=120.0)
    cfg = wl.describe()
    cfg["sample_batch"] = wl.cpu_sample_batch
    line = dict(metric="structured-layer fwd+bwd samples/sec", value=base["value"], unit="samples/s", n_gpus=args.gpus, steps=n,
                warmup=max(args.warmup, 1), ms_per_step=med * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                dtype=wl.dtype, data="synthetic", config=cfg, impl="reference", cpu_baseline=base,
                e2e=dict(value=base["value"], unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_ours(args, wl):
    import torch.distributed as dist
    from structurednets_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    n_gpus = world

    peaks = measured_peaks()
    local_batch = wl.global_batch // n_gpus
    layer = wl.make_layer(device)
    x, gy, labels = wl.make_inputs(local_batch, device, seed=5000 + rank)
    has_flat = hasattr(layer, "flat_grad")
    flat_grad = layer.flat_grad() if has_flat else None
    stream = torch.cuda.current_stream()

    # parameters whose gradient does not live in the flat buffer (sparse / float64 ones): found once, not per step -- walking the
    # 3 501 parameter views of the SSS layer costs more host time than the kernels of a step leave
    loose = [p for p in layer.parameters() if not has_flat or p.is_sparse or p.dtype != torch.float32]

    def zero_grads():
        if has_flat:
            layer.zero_flat_grad()
        for p in loose:
            p.grad = None

    grad_sync = None
    if n_gpus > 1:
        from structurednets_b200.distributed import GradSynchronizer
        grad_sync = GradSynchronizer(layer)

    def sync_grads():
        if grad_sync is not None:
            grad_sync()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(rec=None):
        zero_grads()
        if rec is not None:
            rec[0].record(stream)
        y = layer(x)
        if rec is not None:
            rec[1].record(stream)
        y.backward(gy)
        if rec is not None:
            rec[2].record(stream)
        sync_grads()
        return y

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- pass 1 (eager): warm-up, then K steps with a CUDA-event pair around every library kernel launch (per-kernel durations)
    _lib.lib().sn_timing_enable(1)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    _lib.timing_report()                  # discard the warm-up launches (their events stay in the library's pool)
    _lib.reset_launch_count()
    recs = [(ev(), ev(), ev()) for _ in range(args.steps)]
    t_begin, t_end = ev(), ev()
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_begin.record(stream)
        for i in range(args.steps):
            step(recs[i])
        t_end.record(stream)
        barrier()
    launches = _lib.launch_count()
    _lib.lib().sn_timing_enable(0)
    kernel_ms = _lib.timing_report()      # {kernel: (launches, total ms)} over the eager timed region
    elapsed_ms = t_begin.elapsed_time(t_end)
    eager_elapsed_ms = elapsed_ms
    fwd_ms = float(np.mean([r[0].elapsed_time(r[1]) for r in recs]))
    bwd_ms = float(np.mean([r[1].elapsed_time(r[2]) for r in recs]))

    # ---- pass 2: the same step captured once in a CUDA graph and replayed K times -- launch gaps and host overhead (autograd,
    # ctypes, tensor-map encoding) leave the timed region; the kernels are the same
    graph_note = "eager launches"
    if not args.no_graph:   # multi-GPU: the NCCL all-reduce of the flat gradient is captured with the step
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            g0, g1 = ev(), ev()
            with ClockSampler(local_rank) as clocks:
                barrier()
                g0.record()
                for _ in range(args.steps):
                    graph.replay()
                g1.record()
                barrier()
            elapsed_ms = g0.elapsed_time(g1)
            graph_note = "one step captured in a CUDA graph, replayed %d times" % args.steps
        except Exception as e:   # a layer whose step cannot be captured (host synchronisation inside) keeps the eager number
            torch.cuda.synchronize()
            graph_note = "eager launches (graph capture failed: %s)" % str(e).splitlines()[0][:120]
    if n_gpus > 1:
        t = torch.tensor([elapsed_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = wl.global_batch / (ms_per_step * 1e-3)

    # ---- end to end through the module API with host buffers (pinned H2D + loss + D2H of the loss) ----
    e2e_steps = max(3, min(args.steps, 5))
    x_host = torch.empty((local_batch, wl.input_dim), dtype=x.dtype, pin_memory=True)
    x_host.copy_(x)
    labels_host = labels.cpu().pin_memory()
    loss_fn = torch.nn.CrossEntropyLoss()

    def e2e_step():
        xd = x_host.to(device, non_blocking=True)
        ld = labels_host.to(device, non_blocking=True)
        zero_grads()
        loss = loss_fn(layer(xd).float(), ld)
        loss.backward()
        sync_grads()
        return float(loss.item())     # device -> host read of the step's result

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if n_gpus > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = dict(value=wl.global_batch / e2e_s, unit="samples/s", h2d_bytes_per_step=int(x_host.numel() * x_host.element_size() + labels_host.numel() * 8),
               d2h_bytes_per_step=4, steps=e2e_steps, ms_per_step=e2e_s * 1e3)

    if rank == 0:
        # dominant kernel = the library kernel with the largest share of the timed region; its average launch duration comes from the
        # CUDA-event pairs recorded around every launch (sn_timing_*).  Algorithmic bytes of one launch: the half of SURVEY 8(d)'s
        # per-sample figure that belongs to the pass (forward or backward) the kernel is part of, times the samples of the launch.
        per_kernel = {k: dict(launches=c, avg_ms=t / max(c, 1), share=t / max(eager_elapsed_ms, 1e-9)) for k, (c, t) in kernel_ms.items()}
        dom_name = max(kernel_ms, key=lambda k: kernel_ms[k][1]) if kernel_ms else "n/a"
        dom_ms = per_kernel[dom_name]["avg_ms"] if kernel_ms else max(fwd_ms, bwd_ms)
        achieved = wl.bytes_per_sample_kernel * local_batch / (dom_ms * 1e-3) / 1e9
        roofline = dict(bound=wl.bound, kernel=dom_name, achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s",
                        frac=achieved / peaks["hbm_gbs"], traffic=getattr(wl, "traffic_bytes", {}).get(dom_name), peak_source=peaks["source"],
                        launch_ms=dom_ms, fwd_ms=fwd_ms, bwd_ms=bwd_ms, kernels=per_kernel,
                        step_hbm_frac=wl.bytes_per_sample * local_batch / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        tflops=wl.flop_per_sample * local_batch / (ms_per_step * 1e-3) / 1e12,
                        bf16_tensor_frac_of_measured=(wl.flop_per_sample * local_batch / (ms_per_step * 1e-3) / 1e12 / peaks["bf16_tflops"]) if wl.dtype == "bf16" else None,
                        fp32_fma_tflops=wl.flop_per_sample * local_batch / (ms_per_step * 1e-3) / 1e12,
                        fp32_fma_frac_of_nominal=wl.flop_per_sample * local_batch / (ms_per_step * 1e-3) / 1e12 / FP32_FMA_TFLOPS_NOMINAL)
        cpu_base = None
        if n_gpus == 1 and not args.no_cpu_baseline and wl.cpu_sample_batch:
            cpu_base, _, _ = time_cpu_baseline(wl, steps=20, warmup=2, budget_s=20.0)
        cfg = wl.describe()
        cfg.update(local_batch=local_batch, parallelism="dp%d" % n_gpus, l2_policy="inputs larger than L2 (no flush needed)",
                   launch_mode=graph_note, eager_ms_per_step=eager_elapsed_ms / args.steps)
        line = dict(metric="structured-layer fwd+bwd samples/sec", value=value, unit="samples/s", n_gpus=n_gpus, steps=args.steps,
                    warmup=max(args.warmup, 3), ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype=wl.dtype, data="synthetic", config=cfg, clocks=clocks.summary(), e2e=e2e, gpu_launches=int(launches),
                    roofline=roofline, cpu_baseline=cpu_base)
        print(json.dumps(line))
    if n_gpus > 1:
        # a captured graph holds NCCL work; tearing the process group down underneath it can hang -- make sure every rank is
        # done, flush, and leave without running the destructors
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sss", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--global-batch", type=int, default=None, help="profiling only: override the workload's global batch")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]()
    if args.global_batch:
        wl.global_batch = args.global_batch
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
