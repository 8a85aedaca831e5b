"""Benchmark of the ConvGeodesic + AngularMaxPooling forward+backward hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|tf32|tf32x3|f16x3|ffma] [--impl ours|reference]

Workload (config.workload = "S1"): one FAUST-shaped synthetic mesh per step and per GPU -
V=6890 vertices, 544-dim SHOT-shaped signal, ConvGeodesic 544->96, R5xA8, rotation_delta 1, relu,
followed by AngularMaxPooling; forward + backward (gradients w.r.t. signal, template weights,
self weights, bias) of a random upstream gradient.  Metric: vertices/sec.

One JSON line is printed by rank 0:
  value     device-resident throughput (inputs already in HBM), CUDA-event timed, L2 flushed
            between steps, max over ranks;
  e2e       the same metric through the public layer API starting from pinned HOST buffers
            (signal, bary, upstream grad copied H2D and the pooled output copied D2H every step);
  roofline  the dominant kernel (forward tcgen05 contraction) timed live with CUDA events on its
            launch stream, algorithmic FLOPs / duration against the tensor peak of MEASURED_PEAKS.json: the
            measured bf16 figure for the fp32 mode (fp16 hi/lo operand splits, kind::f16 MMAs, three per
            product: ceiling 1/3, `frac_of_ceiling` reported next to `frac`), bf16/2 (derived; TF32 itself is
            not in that file) for the tf32 modes;
  cpu_baseline  the CPU oracle (a port of the reference algorithm in stock torch CPU ops) timed on
            this host's cores on the same workload.
`--impl reference` times that CPU oracle only (rank 0), with all host threads.
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

S1 = dict(v=6890, f=544, r=5, a=8, t=96, delta=1, act="relu", radius=0.028)
L2_FLUSH_BYTES = 512 << 20


def algorithmic_flops(cfg):
    """fwd 2*V*n_rot*K*T + 2*V*F*T ; bwd (AMP-sparse) 4*V*K*T + 4*V*F*T."""
    v, f, r, a, t = cfg["v"], cfg["f"], cfg["r"], cfg["a"], cfg["t"]
    n_rot = len(range(0, a, cfg["delta"]))
    k = r * a * f
    fwd = 2 * v * n_rot * k * t + 2 * v * f * t
    bwd = 4 * v * k * t + 4 * v * f * t
    return fwd, bwd


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).

    Samples NVML from a thread every 10 ms (the timed region is ~0.1 s: `nvidia-smi -lms` starts too
    slowly to see it); falls back to the nvidia-smi query loop if NVML is unavailable."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.samples = []          # (sm_mhz, reasons_bitmask, power_w)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.gpu_index
            if visible:
                ids = [x for x in visible.split(",") if x.strip()]
                if self.gpu_index < len(ids) and ids[self.gpu_index].strip().isdigit():
                    index = int(ids[self.gpu_index])
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self._nvml = (pynvml, handle)
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
            return
        except Exception:  # noqa: BLE001 - fall back to the command-line tool
            self._nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        pynvml, handle = self._nvml
        while not self._stop.is_set():
            try:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM))
                reasons = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle))
                power = pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0
                self.samples.append((mhz, reasons, power))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        self.samples = []
        self.lines = []

    def stop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            pynvml = self._nvml[0]
            bits = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            sm = sorted(m for m, _r, _p in self.samples)
            reasons = sorted(n for n in names if any(r & bits[n] for _m, r, _p in self.samples))
            power = max((p for _m, _r, p in self.samples), default=None)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "power_w_max": power, "reasons": reasons, "source": "nvml, 10 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 50"}


def make_inputs(cfg, seed_offset=0):
    import torch
    from geoconv_b200 import synthetic as S
    v, f, r, a, t = cfg["v"], cfg["f"], cfg["r"], cfg["a"], cfg["t"]
    signal = S.make_signal(v, f, seed=1 + seed_offset, kind="shot")
    bary = S.make_bary(v, r, a, seed=1 + seed_offset)
    grad_z = torch.randn((v, t), generator=torch.Generator().manual_seed(2 + seed_offset))
    return signal, bary, grad_z


WORKLOAD = "S1: V=6890 F=544 T=96 R5xA8 delta1 relu, ConvGeodesic+AMP fwd+bwd, 1 mesh/step/GPU"


def make_cpu_reference(cfg, weights):
    """The CPU arm: a callable running one fwd+bwd of ConvGeodesic + AngularMaxPooling on the host cores.

    Preferred: the UNMODIFIED reference package (`geoconv`, installed offline into baseline/_ref by
    baseline/install_ref.sh, or an installed `geoconv` in site-packages) through its own public layer API
    -> kind "reference".  If neither is importable: the oracle port (oracle/geoconv_oracle.py) -> kind "port".
    Returns (step(inputs) -> seconds, kind, description)."""
    import torch
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    added = False
    try:
        if os.path.isdir(os.path.join(ref_dir, "geoconv")):
            sys.path.insert(0, ref_dir)
            added = True
        from geoconv.pytorch.layers.angular_max_pooling import AngularMaxPooling as RefAMP
        from geoconv.pytorch.layers.conv_geodesic import ConvGeodesic as RefConv
        import geoconv as _g
        layer = RefConv(input_shape=[(None, cfg["f"]), (None, cfg["r"], cfg["a"], 3, 2)], amt_templates=cfg["t"],
                        template_radius=cfg["radius"], activation=cfg["act"], rotation_delta=cfg["delta"])
        layer.load_state_dict({"_template_neighbor_weights": weights[0], "_template_self_weights": weights[1],
                               "_bias": weights[2]})
        amp = RefAMP()

        def step(inputs):
            signal, bary, grad_z = inputs
            for p_ in layer.parameters():
                p_.grad = None
            xs = signal.clone().requires_grad_(True)
            t0 = time.perf_counter()
            z = amp(layer([xs, bary]))
            z.backward(grad_z)
            return time.perf_counter() - t0
        where = os.path.relpath(os.path.dirname(_g.__file__), ROOT) if _g.__file__.startswith(ROOT) else "site-packages"
        return step, "reference", f"geoconv {where} (unmodified reference layers, torch {torch.__version__} CPU autograd)"
    except Exception as exc:  # noqa: BLE001 - the reference is not reachable: time the port
        if added:
            sys.path.remove(ref_dir)
        sys.stderr.write(f"bench: reference package not importable ({exc}); timing the oracle port\n")
    from oracle import geoconv_oracle as O
    wn, ws, b = weights
    prior = torch.from_numpy(O.geodesic_prior(O.template_matrix(cfg["r"], cfg["a"], cfg["radius"]))).float()

    def step_port(inputs):
        signal, bary, grad_z = inputs
        t0 = time.perf_counter()
        O.conv_amp_forward_backward(signal, bary, wn, ws, b, prior, grad_z, cfg["act"], cfg["delta"])
        return time.perf_counter() - t0
    return step_port, "port", f"oracle/geoconv_oracle.py (port of the reference algorithm, torch {torch.__version__} CPU)"


def run_reference(args):
    import torch
    from geoconv_b200 import synthetic as S
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = S1
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    weights = S.make_weights(cfg["t"], cfg["r"], cfg["a"], cfg["f"], seed=0)
    inputs = make_inputs(cfg)
    step, kind, descr = make_cpu_reference(cfg, weights)
    for _ in range(args.warmup):
        step(inputs)
    total = 0.0
    for _ in range(args.steps):
        total += step(inputs)
    value = cfg["v"] * args.steps / total
    line = {
        "impl": "reference", "metric": "vertices/sec, ConvGeodesic+AMP fwd+bwd", "value": value,
        "unit": "vertices/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "cpu", "implementation": descr},
        "cpu_baseline": {"value": value, "unit": "vertices/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} full S1 meshes (fwd+bwd) after {args.warmup} warm-ups; {descr}"},
        "e2e": {"value": value, "unit": "vertices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))



# ------------------------------------------------------------------------------------------------------------------
# Extras (BASELINE.json configs beyond the headline line): TF32 mode at S1, config 4 (64 meshes, 3-layer net, data
# parallel), config 5 (row-sharded large mesh with halo exchange).  Reported under "extra"; the headline stays S1 fp32.
# ------------------------------------------------------------------------------------------------------------------
def _max_over_ranks(x, world, dev):
    import torch
    import torch.distributed as dist
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def extra_tf32(cfg, weights, inputs_dev, flush, lib, steps, warmup, peaks):
    """S1 in the TF32 mode (<= 2e-3): device-resident step from a CUDA graph, forward kernel timed with events."""
    import ctypes
    import torch
    from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
    from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
    signal_d, bary_d, gz_d = inputs_dev
    dev = signal_d.device
    v, f, r, a, t = cfg["v"], cfg["f"], cfg["r"], cfg["a"], cfg["t"]
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=cfg["radius"],
                         activation=cfg["act"], rotation_delta=cfg["delta"], precision="tf32")
    layer.load_state_dict({"_template_neighbor_weights": weights[0], "_template_self_weights": weights[1],
                           "_bias": weights[2]})
    layer = layer.to(dev)
    amp = AngularMaxPooling()
    xs = signal_d.detach().clone().requires_grad_(True)

    def body():
        for p_ in layer.parameters():
            p_.grad = None
        xs.grad = None
        amp(layer([xs, bary_d])).backward(gz_d)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body()
    for _ in range(warmup):
        flush.zero_()
        graph.replay()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a_.elapsed_time(b_) for a_, b_ in evs) / steps
    lib.gcb_profile_enable(1)
    for _ in range(min(steps, 10)):
        flush.zero_()
        body()
    torch.cuda.synchronize()
    k_ms, k_cnt = ctypes.c_double(0.0), ctypes.c_int64(0)
    lib.gcb_profile_read(0, ctypes.byref(k_ms), ctypes.byref(k_cnt))
    lib.gcb_profile_enable(0)
    fwd_flops, _bwd = algorithmic_flops(cfg)
    peak = peaks["bf16_tflops"] / 2.0
    achieved = fwd_flops * k_cnt.value / (k_ms.value * 1e-3) / 1e12 if k_ms.value > 0 else None
    return {"workload": WORKLOAD, "precision": "tf32 (single pass, <= 2e-3 normalised error)", "ms_per_step": ms,
            "value": v / (ms * 1e-3), "unit": "vertices/s", "launch": "cuda-graph replay",
            "roofline": {"bound": "tensor", "kernel": "k_tc_conv", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if achieved else None,
                         "kernel_ms_avg": k_ms.value / k_cnt.value if k_cnt.value else None,
                         "peak_source": f"{peaks['source']} bf16_tflops/2 (TF32 dense peak is derived, not measured)"}}


def extra_s3(rank, world, dev, steps=3, warmup=1, n_meshes=64):
    """BASELINE config 4: 3-layer ConvGeodesic + AngularMaxPooling segmentation net (64 -> 96 -> 256 -> 384, Linear head,
    cross entropy) on a batch of 64 FAUST-shaped meshes, meshes dealt round-robin to the ranks, one all-reduce of the
    accumulated weight gradients per step.  Also: the data-parallel gradients against the same batch on one rank."""
    import torch
    import torch.distributed as dist
    from geoconv_b200 import synthetic as S
    from geoconv_b200.distributed import allreduce_gradients
    from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
    from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
    v, r, a, n_cls = 6890, 5, 8, 32
    dims = [64, 96, 256, 384]

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.convs = torch.nn.ModuleList(
                ConvGeodesic(input_shape=[(None, fi), (None, r, a, 3, 2)], amt_templates=fo, template_radius=0.028)
                for fi, fo in zip(dims[:-1], dims[1:]))
            self.amp = AngularMaxPooling()
            self.head = torch.nn.Linear(dims[-1], n_cls)

        def forward(self, x, bary):
            for conv in self.convs:
                x = self.amp(conv([x, bary]))
            return self.head(x)
    torch.manual_seed(0)
    net = Net().to(dev)
    params = list(net.parameters())
    loss_fn = torch.nn.CrossEntropyLoss(reduction="sum")
    mine = list(range(rank, n_meshes, world))
    meshes = {}
    for m in (range(n_meshes) if world > 1 and n_meshes <= 64 else mine):
        g = torch.Generator().manual_seed(1000 + m)
        meshes[m] = (S.make_signal(v, dims[0], seed=100 + m).to(dev), S.make_bary(v, r, a, seed=100 + m).to(dev),
                     torch.randint(0, n_cls, (v,), generator=g).to(dev))

    # This rank's meshes as vertex-concatenated batches (block-diagonal: the barycentric indices of mesh k are offset
    # by k * V - the layer API takes any vertex count): one layer call per batch instead of one per mesh, so the
    # per-call work (prior folding, operand maxima and splits, ~25 launches per layer) is paid once per batch and the
    # kernels see 8 x as many vertex tiles.  GCB_S3_CHUNK=1: one call per mesh (A/B).
    chunk = max(1, int(os.environ.get("GCB_S3_CHUNK", "8")))

    def make_batches(ids_all):
        out = []
        for i in range(0, len(ids_all), chunk):
            ids = ids_all[i:i + chunk]
            barys = []
            for k, m in enumerate(ids):
                b = meshes[m][1].clone()
                b[..., 0] += k * v
                barys.append(b)
            out.append((torch.cat([meshes[m][0] for m in ids]), torch.cat(barys), torch.cat([meshes[m][2] for m in ids])))
        return out
    batches = make_batches(mine)

    def step():
        for p_ in params:
            p_.grad = None
        for x, bary, y in batches:
            loss_fn(net(x, bary), y).backward()
        if world > 1:
            allreduce_gradients(params)
    grad_err = None
    step()
    if world > 1:
        dp = [p_.grad.clone() for p_ in params]
        # the whole batch on this rank, in the same vertex-concatenated batches the ranks used (same arithmetic, so
        # what remains is the summation order of the all-reduce)
        for p_ in params:
            p_.grad = None
        for rr in range(world):
            for x, bary, y in make_batches(list(range(rr, n_meshes, world))):
                loss_fn(net(x, bary), y).backward()
        err = max(((x_.double() - p_.grad.double()).abs().max() / p_.grad.double().abs().max().clamp_min(1e-30)).item()
                  for x_, p_ in zip(dp, params))
        grad_err = _max_over_ranks(err, world, dev)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, world, dev)
    return {"workload": f"S3 (BASELINE config 4): {n_meshes} FAUST-shaped meshes (V=6890, R5xA8), net 64->96->256->384 + "
                        f"Linear->{n_cls}, cross entropy, fwd+bwd, fp32 mode, meshes round-robin over {world} rank(s), "
                        f"in vertex-concatenated batches of {chunk}, one all-reduce of the accumulated weight gradients per step",
            "ms_per_step": ms, "value": n_meshes * v / (ms * 1e-3), "unit": "vertices/s", "steps": steps, "launch": "eager",
            "grad_max_normalised_err_vs_one_rank": grad_err}


def extra_s4(rank, world, dev, rows_per_rank=250_000, iters=5):
    """BASELINE config 5 (at N = 8: one mesh of 2 M vertices): ConvGeodesic 128 -> 128 forward on a mesh partitioned by
    vertex rows, 250 k rows per rank, halo exchange of neighbour signal rows over NCCL, with and without overlapping
    the exchange with the interior rows.  Every rank generates its own rows only (locality-preserving indices)."""
    import torch
    import torch.distributed as dist
    from geoconv_b200 import synthetic as S
    from geoconv_b200.distributed import ShardedForward, build_local_shard_plan, row_ranges
    from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
    f = t = 128
    r, a = 5, 8
    v_total = rows_per_rank * world
    lo, hi = row_ranges(v_total, world)[rank]
    bary_rows = S.make_bary_local_rows(v_total, lo, hi, r, a, window=2048, seed=5)
    plan = build_local_shard_plan(bary_rows, v_total)
    torch.manual_seed(1)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t, template_radius=0.028).to(dev)
    sf = ShardedForward(layer, plan, f, dev)
    sf.load(S.make_signal(hi - lo, f, seed=50 + rank, kind="randn").to(dev))
    out = {}
    with torch.no_grad():
        ys = {}
        for overlap in (True, False):
            for _ in range(2):
                sf(overlap=overlap)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                blocks = sf(overlap=overlap)
            e1.record()
            torch.cuda.synchronize()
            out[overlap] = _max_over_ranks(e0.elapsed_time(e1) / iters, world, dev)
            ys[overlap] = [b_ for b_ in blocks if b_ is not None]
        same = all(torch.equal(x_, y_) for x_, y_ in zip(ys[True], ys[False]))
    ok = _max_over_ranks(0.0 if same else 1.0, world, dev) == 0.0
    return {"workload": f"S4 (BASELINE config 5 at N = 8): one mesh of {v_total} vertices, ConvGeodesic 128->128 R5xA8 forward, "
                        f"fp32 mode, {rows_per_rank} rows per rank, halo exchange (all_to_all of packed signal rows) over NCCL",
            "ms_forward_overlapped": out[True], "ms_forward_serial": out[False],
            "value": v_total / (out[True] * 1e-3), "unit": "vertices/s",
            "rows_interior": sf.hi - sf.lo, "rows_boundary": plan.n_local - (sf.hi - sf.lo), "halo_rows": plan.n_halo,
            "halo_bytes_per_rank": sf.halo_bytes, "overlap_equals_serial": ok}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from geoconv_b200 import _native, ops
    from geoconv_b200.distributed import allreduce_gradients, overlapped_grad_sync
    from geoconv_b200.pytorch.layers.angular_max_pooling import AngularMaxPooling
    from geoconv_b200.pytorch.layers.conv_geodesic import ConvGeodesic
    from geoconv_b200 import synthetic as S   # the oracle is imported by the cpu_baseline leg only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = _native.load()
    cfg = S1
    v, f, r, a, t = cfg["v"], cfg["f"], cfg["r"], cfg["a"], cfg["t"]

    weights = S.make_weights(t, r, a, f, seed=0)
    layer = ConvGeodesic(input_shape=[(None, f), (None, r, a, 3, 2)], amt_templates=t,
                         template_radius=cfg["radius"], activation=cfg["act"], rotation_delta=cfg["delta"],
                         precision=args.precision)
    layer.load_state_dict({"_template_neighbor_weights": weights[0], "_template_self_weights": weights[1],
                           "_bias": weights[2]})
    layer = layer.to(dev)
    amp = AngularMaxPooling()
    # The step's three inputs live in ONE pinned, write-combined host arena and travel in ONE copy
    # (geoconv_b200/data/staging.py: a pinned buffer the CPU has just written is copied at ~25 GB/s on this box - every
    # DMA read snoops the dirty cache lines - write-combined pinned memory at ~45 GB/s).
    from geoconv_b200.data.staging import PinnedArena, device_arena
    inputs = make_inputs(cfg, seed_offset=rank)
    host_arena = PinnedArena(inputs).fill(inputs)
    arena_h = host_arena.buffer
    signal_h, bary_h, gz_h = inputs          # (CPU copies for the resident leg; the arena is write-only for the CPU)

    def arena(tensors, device=None):
        return device_arena(tensors, device)
    signal_d = signal_h.to(dev).requires_grad_(True)
    bary_d = bary_h.to(dev)
    gz_d = gz_h.to(dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    params = [layer._template_neighbor_weights, layer._template_self_weights, layer._bias]

    # Data parallel over meshes (N > 1): the weight gradients are summed over the ranks.  Preferred: the all-reduce
    # is started INSIDE the layer's backward, right after the weight-gradient stage, and runs beside the dSignal
    # GEMM (geoconv_b200.distributed.overlapped_grad_sync); it is then part of the captured step.  If that cannot
    # be captured in a CUDA graph the step is replayed without it and the gradients are reduced after the replay.
    grad_sync = {"ctx": None, "mode": "none"}
    if world > 1 and args.overlap_allreduce:
        grad_sync["ctx"] = overlapped_grad_sync(max_ctas=args.allreduce_ctas, peer=bool(args.peer_allreduce))
        grad_sync["ctx"].__enter__()
        grad_sync["mode"] = "all-reduce inside the backward (after the weight-gradient stage, beside the dSignal GEMM)"
    elif world > 1:
        grad_sync["mode"] = "all-reduce after the step"

    def reduce_after_step():
        if world > 1 and grad_sync["ctx"] is None:
            allreduce_gradients(params)

    def drop_in_backward_sync(why):
        if grad_sync["ctx"] is not None:
            grad_sync["ctx"].__exit__(None, None, None)
            grad_sync["ctx"] = None
            grad_sync["mode"] = f"all-reduce after the step ({why})"

    def step_resident():
        for p_ in params:
            p_.grad = None
        signal_d.grad = None
        z = amp(layer([signal_d, bary_d]))
        z.backward(gz_d)
        reduce_after_step()
        return z

    # buffers for the host-to-host leg (one device arena, the three tensors are views of it)
    arena_buf, (sig_buf, bary_buf, gz_buf) = arena(inputs, device=dev)
    z_host = torch.empty((v, t), dtype=torch.float32).pin_memory()

    def step_e2e():
        for p_ in params:
            p_.grad = None
        arena_buf.copy_(arena_h, non_blocking=True)   # bumps the views' version: bary is re-packed, CSR rebuilt
        xs = sig_buf.detach().requires_grad_(True)
        z = amp(layer([xs, bary_buf]))
        z.backward(gz_buf)
        reduce_after_step()
        z_host.copy_(z.detach(), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, profile=False, sampler=None):
        for _ in range(warmup):
            flush.zero_()
            step_fn()
        barrier()
        if sampler is not None:
            sampler.mark()
        if profile:
            lib.gcb_profile_enable(1)
        launches0 = lib.gcb_launch_count()
        evs = []
        for _ in range(steps):
            flush.zero_()                     # evict the previous step's working set from L2
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        total_ms = sum(a_.elapsed_time(b_) for a_, b_ in evs)
        launches = lib.gcb_launch_count() - launches0
        if world > 1:
            tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            total_ms = tt.item()
        return total_ms, launches

    # Device-resident leg: the whole fwd+bwd step (22 kernels) is captured once in a CUDA graph and
    # replayed, so launch latency of the short kernels is off the critical path.  The weight-gradient
    # all-reduce (N > 1) stays outside the graph.  Falls back to eager launches if capture fails.
    launch_mode = "eager"
    step_value = step_resident
    launches_per_step = None
    if args.graph:
        try:
            def fwd_bwd():
                for p_ in params:
                    p_.grad = None
                signal_d.grad = None
                z = amp(layer([signal_d, bary_d]))
                z.backward(gz_d)
                return z
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fwd_bwd()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = lib.gcb_launch_count()
            with torch.cuda.graph(graph):
                fwd_bwd()
            launches_per_step = lib.gcb_launch_count() - l0
            graph.replay()
            torch.cuda.synchronize()

            def step_graph():
                graph.replay()
                reduce_after_step()
            step_value = step_graph
            launch_mode = "cuda-graph replay"
        except Exception as exc:  # noqa: BLE001 - measurement falls back, it never hides a compute error
            sys.stderr.write(f"bench: CUDA graph capture failed ({exc}); timing eager launches\n")
            torch.cuda.synchronize()
            if grad_sync["ctx"] is not None:
                # the collective inside the backward could not be captured: same capture without it
                drop_in_backward_sync("a collective inside the backward could not be captured")
                try:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        fwd_bwd()
                    graph.replay()
                    torch.cuda.synchronize()

                    def step_graph2():
                        graph.replay()
                        reduce_after_step()
                    step_value = step_graph2
                    launch_mode = "cuda-graph replay"
                except Exception as exc2:  # noqa: BLE001
                    sys.stderr.write(f"bench: second capture failed too ({exc2}); eager launches\n")
                    torch.cuda.synchronize()

    total_ms, launches = timed(step_value, args.steps, args.warmup, sampler=sampler)
    if launches_per_step is not None:
        launches = launches_per_step * args.steps
    # roofline leg: the same steps launched eagerly with CUDA events around the dominant kernel
    timed(step_resident, args.steps, 1, profile=True)
    import ctypes
    prof = {}
    for which, name in ((0, "fwd_conv"),):
        ms = ctypes.c_double(0.0)
        cnt = ctypes.c_int64(0)
        lib.gcb_profile_read(which, ctypes.byref(ms), ctypes.byref(cnt))
        prof[name] = (ms.value, cnt.value)
    lib.gcb_profile_enable(0)
    clocks = sampler.stop()
    def timed_pipelined(steps, warmup):
        """Host-to-host steps with the next step's H2D overlapped with this step's compute (what a prefetching
        input pipeline does).  Two buffer sets, one CUDA graph of the user's step per set (bary split, CSR
        inversion, fwd, bwd through the public layer API), copies issued eagerly on a copy stream.
        Returns the total time of `steps` steps with the time of the `steps` L2 flushes subtracted."""
        from geoconv_b200 import ops
        prefetch_prep = os.environ.get("GCB_BENCH_PREFETCH_PREP", "1") != "0"   # A/B
        main = torch.cuda.current_stream()
        copy = torch.cuda.Stream()
        back = torch.cuda.Stream()   # results go home on their own stream (and copy engine): a D2H queued between two
                                     # H2D copies made the next step's inputs wait for this step's result
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)   # the parameters' AccumulateGrad nodes were created on another stream by the earlier legs
        sets = []
        for _ in range(2):
            a_, (s_, b_, g_) = arena(inputs, device=dev)
            bufs = dict(arena=a_, sig=s_, bary=b_, gz=g_)
            bufs["arena"].copy_(arena_h)

            def body(bufs=bufs):
                for p_ in params:
                    p_.grad = None
                xs = bufs["sig"].detach().requires_grad_(True)
                z = amp(layer([xs, bufs["bary"]]))
                z.backward(bufs["gz"])
                return z

            def prepare(bufs=bufs):
                # the per-mesh preparation through the public API: barycentric split (index validation on), gather
                # records, CSR inversion - every step, because every step brings a new barycentric tensor
                pack = ops.prep_bary(bufs["bary"], bufs["sig"].shape[0])
                pack.records()
                pack.csr()
                return pack
            side = torch.cuda.Stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for _ in range(2):
                    bufs["arena"].copy_(arena_h)        # new version: the step re-packs the barycentric tensor
                    if prefetch_prep:
                        ops.attach_pack(bufs["bary"], prepare())
                    body()
            main.wait_stream(side)
            torch.cuda.synchronize()
            bufs["arena"].copy_(arena_h)
            torch.cuda.synchronize()
            if prefetch_prep:
                # stage 1 of the pipeline (prefetch stream, behind the H2D copy): the preparation, its own graph
                gp = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gp):
                    bufs["pack"] = prepare()
                bufs["prep_graph"] = gp
                ops.attach_pack(bufs["bary"], bufs["pack"])
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                bufs["z"] = body()
            bufs["graph"] = g
            bufs["grads"] = [p_.grad for p_ in params]
            sets.append(bufs)
        torch.cuda.synchronize()

        def h2d(bufs):
            bufs["arena"].copy_(arena_h, non_blocking=True)
            if prefetch_prep:
                bufs["prep_graph"].replay()          # (on the prefetch stream: beside the previous step's compute)

        def run(n):
            done = [None, None]                      # compute-finished event of the last step on each set
            with torch.cuda.stream(copy):            # step 0's inputs: nothing to hide them behind
                copy.wait_stream(main)
                h2d(sets[0])
                copied = torch.cuda.Event()
                copied.record(copy)
            for i in range(n):
                cur, nxt = sets[i & 1], sets[(i + 1) & 1]
                flush.zero_()
                flushed = torch.cuda.Event()
                flushed.record(main)
                copied_next = None
                if i + 1 < n:
                    with torch.cuda.stream(copy):
                        copy.wait_event(flushed)
                        if done[(i + 1) & 1] is not None:
                            copy.wait_event(done[(i + 1) & 1])   # the set is free once its previous step has finished
                        h2d(nxt)
                        copied_next = torch.cuda.Event()
                        copied_next.record(copy)
                main.wait_event(copied)
                cur["graph"].replay()
                for p_, g_ in zip(params, cur["grads"]):
                    p_.grad = g_
                reduce_after_step()
                fin = torch.cuda.Event()
                fin.record(main)
                done[i & 1] = fin
                with torch.cuda.stream(back):
                    back.wait_event(fin)
                    z_host.copy_(cur["z"].detach(), non_blocking=True)
                    read = torch.cuda.Event()
                    read.record(back)
                done[i & 1] = read                   # the set (its z included) is free once the result has left
                copied = copied_next
            main.wait_stream(copy)
            main.wait_stream(back)

        run(warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        barrier()
        total = e0.elapsed_time(e1)
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(steps):
            flush.zero_()
        f1.record()
        barrier()
        total -= f0.elapsed_time(f1)
        if world > 1:
            tt = torch.tensor([total], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            total = tt.item()
        return total

    # Host-to-host leg: the same public layer calls, with the H2D copies of this step's inputs, the
    # per-mesh preparation (bary split, CSR inversion) and the D2H of the result inside the timed
    # region.  The user's step is captured in a CUDA graph with the layer's defaults (index validation is
    # a device-side counter inside gcb_prep_bary, no host sync); eager launches if capture fails.
    e2e_mode = "eager"
    step_e2e_value = step_e2e
    if args.graph:
        try:
            def e2e_body():   # (index validation stays on: it is a device-side counter, not a host sync)
                for p_ in params:
                    p_.grad = None
                arena_buf.copy_(arena_h, non_blocking=True)
                xs = sig_buf.detach().requires_grad_(True)
                z = amp(layer([xs, bary_buf]))
                z.backward(gz_buf)
                z_host.copy_(z.detach(), non_blocking=True)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    e2e_body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph2):
                e2e_body()
            graph2.replay()
            torch.cuda.synchronize()

            def step_e2e_graph():
                graph2.replay()
                reduce_after_step()
            step_e2e_value = step_e2e_graph
            e2e_mode = "cuda-graph replay"
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write(f"bench: e2e CUDA graph capture failed ({exc}); timing eager launches\n")
            torch.cuda.synchronize()
    e2e_serial_ms, _ = timed(step_e2e_value, args.steps, args.warmup)
    e2e_ms, e2e_mode_p = e2e_serial_ms, None
    if args.e2e_pipeline and args.graph:
        try:
            e2e_ms = timed_pipelined(args.steps, args.warmup)
            torch.cuda.synchronize()
            # the pipelined steps must deliver the same result to the host as the resident step
            z_ref = step_resident().detach().cpu()
            if not torch.allclose(z_host, z_ref, rtol=0, atol=1e-6 * float(z_ref.abs().max())):
                raise RuntimeError("pipelined host-to-host step returned a different pooled output")
            e2e_mode_p = ("pipelined: while step n computes, the inputs of step n+1 are copied H2D on a second stream into "
                          "the other of two buffer sets (never during the L2 flush)"
                          + (" and prepared there (barycentric split, gather records, CSR inversion: every step)"
                             if os.environ.get("GCB_BENCH_PREFETCH_PREP", "1") != "0" else "")
                          + "; every step's H2D, preparation and D2H are inside the timed region, the first copy exposed")
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write(f"bench: pipelined e2e failed ({exc}); reporting the serial number\n")
            torch.cuda.synchronize()
            e2e_ms = e2e_serial_ms

    if grad_sync["ctx"] is not None:
        sync_obj = grad_sync["ctx"].sync
        if sync_obj.peer_reducer is not None and sync_obj.peer_reducer.calls:
            grad_sync["mode"] += ("; collective = gcb_peer_allreduce_staged, the library's own two-shot all-reduce kernel on "
                                  "NVLink peer memory (12 CTAs beside the dSignal GEMM)")
        else:
            grad_sync["mode"] += f"; collective = NCCL on a low-CTA high-priority communicator ({sync_obj.peer_failed or 'peer kernel not used'})"
    extra = {}
    if args.extras:
        if grad_sync["ctx"] is not None:       # the extras synchronise their gradients themselves
            grad_sync["ctx"].__exit__(None, None, None)
        for name, fn in (("tf32", (lambda: extra_tf32(cfg, weights, (signal_d, bary_d, gz_d), flush, lib, args.steps,
                                                        args.warmup, load_peaks())) if world == 1 else None),
                         ("s3", lambda: extra_s3(rank, world, dev)),
                         ("s4", (lambda: extra_s4(rank, world, dev)) if world > 1 else None)):
            if fn is None:
                continue
            try:
                extra[name] = fn()
            except Exception as exc:  # noqa: BLE001 - an extra never takes the headline line down
                extra[name] = {"error": f"{type(exc).__name__}: {exc}"}
                torch.cuda.synchronize()

    def leave_multi_rank():
        # CUDA graphs that hold NCCL kernels are still alive here; tearing the communicator down under them made
        # rank 0 hang in destroy_process_group (N = 2, r02).  Nothing is left to do collectively: flush and leave.
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)

    if rank != 0:
        if world > 1:
            leave_multi_rank()
        return

    value = world * v * args.steps / (total_ms * 1e-3)
    e2e_value = world * v * args.steps / (e2e_ms * 1e-3)
    fwd_flops, bwd_flops = algorithmic_flops(cfg)
    peaks = load_peaks()
    # fp32 mode contracts fp16 hi/lo operand splits with kind::f16 MMAs (same rate as bf16: the MEASURED peak),
    # three per product; tf32 mode (and GEOCONV_B200_FP32_VIA=tf32x3) runs kind::tf32 at half that rate (derived)
    fp32_via_f16 = args.precision == "f16x3" or (args.precision == "fp32" and ops.FP32_VIA_F16)
    three_products = args.precision in ("fp32", "f16x3", "tf32x3")
    tensor_peak = peaks["bf16_tflops"] if fp32_via_f16 else peaks["bf16_tflops"] / 2.0
    ceiling = (1.0 / 3.0) if three_products else 1.0
    k_ms, k_cnt = prof["fwd_conv"]
    achieved = (fwd_flops * k_cnt / (k_ms * 1e-3)) / 1e12 if k_ms > 0 and k_cnt else None
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):      # DRAM bytes of one launch, from the committed ncu --set full capture
        with open(tpath) as fh:
            rec = json.load(fh).get(args.precision)
        if rec:
            traffic = rec["dram_bytes_read"] + rec["dram_bytes_write"]
            traffic_src = rec["source"]
    roofline = {
        "bound": "tensor", "kernel": "k_tc_conv (forward gather+tcgen05 contraction)",
        "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
        "frac": (achieved / tensor_peak) if achieved else None,
        "ceiling_frac": ceiling, "frac_of_ceiling": (achieved / (tensor_peak * ceiling)) if achieved else None,
        "traffic": traffic,
        "traffic_unit": "bytes of DRAM read+write per launch (ncu)", "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": 4 * (v * f + 6 * v * r * a + t * (r * a * f + f) + t + v * t) + 4 * v,
        "peak_source": (f"{peaks['source']} bf16_tflops (kind::f16 issues at the bf16 rate)" if fp32_via_f16 else
                        f"{peaks['source']} bf16_tflops/2 (TF32 dense peak is derived, not measured)"),
        "algorithmic_flops_per_launch": fwd_flops, "kernel_ms_avg": (k_ms / k_cnt) if k_cnt else None,
        "note": (("fp32-faithful mode: operands split into fp16 hi + lo (power-of-two scaled), 3 kind::f16 MMAs per "
                  "product, so algorithmic FLOP/s cannot exceed 1/3 of this peak" if fp32_via_f16 else
                  "fp32-faithful mode: 3 tf32 MMAs per product (3xTF32): its ceiling against this peak is 1/3")
                 if three_products else "single-pass tf32"),
    }

    # CPU baseline: the oracle on this host's cores, bounded sample (1 warm-up + 2 timed meshes)
    # (timed at N = 1 only; multi-rank runs report null and rely on `--impl reference`)
    cpu_baseline = None
    if world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_inputs = make_inputs(cfg)
        cpu_step, cpu_kind, cpu_descr = make_cpu_reference(cfg, weights)
        cpu_step(cpu_inputs)
        cpu_t = min(cpu_step(cpu_inputs) for _ in range(2))
        cpu_baseline = {"value": v / cpu_t, "unit": "vertices/s", "cores": cores, "kind": cpu_kind,
                        "sample": f"best of 2 full S1 meshes (fwd+bwd) after 1 warm-up; {cpu_descr}"}

    h2d = arena_h.numel()   # signal + barycentric tensor + upstream gradient, padded to 256 bytes each: one copy
    d2h = z_host.numel() * 4
    line = {
        "metric": "vertices/sec, ConvGeodesic+AMP fwd+bwd", "value": value, "unit": "vertices/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32" if args.precision == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "precision": args.precision, "launch": launch_mode, "l2": f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB write)",
                   "parallelism": f"dp{world} (one mesh per rank per step" + (f", NCCL {grad_sync['mode']})" if world > 1 else ")")},
        "e2e": {"value": e2e_value, "unit": "vertices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "launch": e2e_mode,
                "overlap": e2e_mode_p or "none (copies and compute serialised on one stream)",
                "serial_ms_per_step": e2e_serial_ms / args.steps,
                "h2d_gbs_per_rank": h2d / (e2e_ms / args.steps * 1e-3) / 1e9,
                "h2d_gbs_all_ranks": world * h2d / (e2e_ms / args.steps * 1e-3) / 1e9,
                "h2d_gbs_note": "input bytes per step / step time = the rate the host has to sustain, not the speed of the copy "
                                "(one 24 MB copy from the write-combined arena runs at ~45 GB/s and is hidden behind compute)",
                "derivation": ("pipelined figure = device time of the timed loop minus the separately timed loop of its L2 "
                               "flushes (the flush is measurement apparatus, not part of a step); serial figure = per-step "
                               "events, flush outside") if e2e_mode_p else "per-step events, flush outside",
                "index_validation": "on (layer default: device-side counter, deferred report)",
                "includes": "H2D signal+bary+upstream grad (one pinned write-combined arena, one copy per step), bary split, CSR inversion, fwd, bwd, D2H pooled output"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
        "flops_per_step": {"fwd": fwd_flops, "bwd": bwd_flops},
        "extra": extra,
    }
    emit(json.dumps(line))
    if world > 1:
        leave_multi_rank()


_RESULT_FD = None


def emit(text):
    """The one JSON line, written to the process's ORIGINAL stdout."""
    if _RESULT_FD is None:
        print(text, flush=True)
    else:
        os.write(_RESULT_FD, (text + "\n").encode())


def quiet_stdout():
    """stdout carries exactly one JSON line: anything a library prints there at the C level (NCCL writes
    "NCCL version ..." to stdout when NCCL_DEBUG is set in the environment) is sent to stderr instead."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32", "ffma", "tf32x3", "f16x3"])
    ap.add_argument("--graph", type=int, default=1, help="1: replay the resident step from a CUDA graph (default)")
    ap.add_argument("--overlap-allreduce", type=int, default=1,
                    help="1: N > 1 starts the weight-gradient all-reduce inside the backward (default); 0: after the step")
    ap.add_argument("--peer-allreduce", type=int, default=1,
                    help="1: the in-backward all-reduce is the library's own NVLink peer-memory kernel (default); 0: NCCL")
    ap.add_argument("--allreduce-ctas", type=int, default=8, help="CTA cap of the in-backward all-reduce's NCCL communicator")
    ap.add_argument("--extras", type=int, default=1, help="1: also measure the TF32 mode (N = 1) and configs 4 / 5 (default)")
    ap.add_argument("--e2e-pipeline", type=int, default=1,
                    help="1: overlap the H2D of step n+1 with the compute of step n in the host-to-host leg (default)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
