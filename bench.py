#!/usr/bin/env python
"""bench.py -- headline benchmark (see BASELINE.json): MB/s of encode+decode, HBM-roofline fraction, CPU reference beside it.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU implementation on the host cores

Workload at N=1 = BASELINE.json configs[1]: LZNT1 encode+decode of a 1 GiB synthetic text stream cut into 4 KiB chunks
(1024 x 1 MiB blocks of generator `text`, seeds 42+i).  One "step" = one encode of the whole stream plus one
decode of the result, device-resident (inputs already in HBM when the timed region starts).  With N > 1 every rank holds its
own 1 GiB shard (weak scaling; units are independent) and the only exchange is one all-gather of the per-rank
output sizes per step.  `e2e` is the same round trip through the C ABI (lznt1_compress / lznt1_decompress) with pinned HOST
buffers, host<->device copies inside the timed region.  Extra keys report the other formats once their kernels are linked.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MB = 1e6
BLOCK = 1 << 20
METRIC = "LZNT1 encode+decode throughput (uncompressed MB/s, 1 GiB synthetic text stream as 4 KiB chunks per GPU)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref, compiled from the unmodified sources) or, if that file is
# absent, the oracle port -- timed on the host cores, one process per core (threads do not scale).
# ------------------------------------------------------------------------------------------------------------------
def _cpu_lib():
    ref = os.path.join(ROOT, "oracle", "_ref", "libMSCompression.so")
    if os.path.exists(ref):
        return ref, "", "reference"
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    return os.path.join(ROOT, "oracle", "liboracle.so"), "orc_", "port"


def host_cores():
    """Host threads this process may really use: the scheduler affinity mask, clipped by the cgroup CPU quota if there is one."""
    aff = len(os.sched_getaffinity(0))
    quota = None
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            q, per = open(path).read().split()[:2]
            if q != "max":
                quota = float(q) / float(per)
        except Exception:
            pass
    if quota is None:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    usable = aff if quota is None else max(1, min(aff, int(quota + 0.5)))
    return usable, aff, quota


CPU_UNIT = {"lznt1": BLOCK, "xpress": BLOCK, "xpress_huff": 65536}       # bytes per call: what the GPU arm's batch unit is


def _cpu_proc(path, prefix, fmt, seed0, nunits, reps, barrier, queue):
    """One worker: generate its units, warm up, wait for everybody, then time compress+decompress of every unit."""
    import mscompress_b200 as ms
    lib = ctypes.CDLL(path)
    comp_f, dec_f = getattr(lib, prefix + fmt + "_compress"), getattr(lib, prefix + fmt + "_decompress")
    unit = CPU_UNIT[fmt]
    blocks = [ms.generate("text", unit, seed0 + i) for i in range(nunits)]
    cap = unit + unit // 8 + 4096
    cbuf, dbuf = ctypes.create_string_buffer(cap), ctypes.create_string_buffer(unit)
    n = ctypes.c_size_t(cap)
    comp_f(blocks[0], ctypes.c_size_t(unit), cbuf, ctypes.byref(n))              # warm (code + buffers paged in)
    t_c = t_d = 0.0
    csize = 0
    barrier.wait()
    t_begin = time.perf_counter()                                                # CLOCK_MONOTONIC: comparable across processes
    for _ in range(reps):
        for b in blocks:
            n = ctypes.c_size_t(cap)
            t0 = time.perf_counter()
            rc = comp_f(b, ctypes.c_size_t(unit), cbuf, ctypes.byref(n))
            t1 = time.perf_counter()
            m = ctypes.c_size_t(unit)
            rc2 = dec_f(cbuf, n, dbuf, ctypes.byref(m))
            t2 = time.perf_counter()
            assert rc == 0 and rc2 == 0 and m.value == unit
            t_c += t1 - t0
            t_d += t2 - t1
            csize += n.value
    queue.put((nunits * reps * unit, t_c, t_d, csize, t_begin, time.perf_counter()))


def cpu_measure(fmt="lznt1", bytes_per_proc=8 * BLOCK, reps=1, procs=None):
    """Aggregate MB/s of compress+decompress of text units (1 MiB; 64 KiB for Xpress-Huffman) over `procs` worker processes.
    Data generation and warm-up happen before a barrier; the wall clock runs from the first worker's start to the last one's end."""
    import multiprocessing as mp
    path, prefix, kind = _cpu_lib()
    usable, aff, quota = host_cores()
    procs = procs or usable
    nunits = max(1, bytes_per_proc // CPU_UNIT[fmt])
    ctx = mp.get_context("fork")
    barrier, queue = ctx.Barrier(procs), ctx.Queue()
    ps = [ctx.Process(target=_cpu_proc, args=(path, prefix, fmt, 42 + p * nunits, nunits, reps, barrier, queue)) for p in range(procs)]
    for p in ps:
        p.start()
    res = [queue.get(timeout=600) for _ in ps]
    for p in ps:
        p.join()
    wall = max(r[5] for r in res) - min(r[4] for r in res)
    total = sum(r[0] for r in res)
    tc, td = sum(r[1] for r in res), sum(r[2] for r in res)
    return {
        "value": total / MB / wall, "unit": "MB/s", "cores": procs, "kind": kind,
        "sample": "%d processes x %d x %d KiB text units, %s compress+decompress via %s (generation and warm-up outside the timed wall)"
                  % (procs, nunits * reps, CPU_UNIT[fmt] // 1024, fmt, os.path.basename(path)),
        "compress_MBps_per_core": total / MB / tc, "decompress_MBps_per_core": total / MB / td,
        "ratio": sum(r[3] for r in res) / total, "wall_s": wall,
        "host": {"affinity_cpus": aff, "cgroup_cpu_quota": quota, "os_cpu_count": os.cpu_count()},
    }


def cpu_baseline_block(fmt, bytes_per_proc):
    """All-core and one-core figures of one format, in the shape the JSON line carries."""
    full = cpu_measure(fmt, bytes_per_proc)
    one = cpu_measure(fmt, bytes_per_proc, procs=1)
    return {"value": full["value"], "unit": "MB/s", "cores": full["cores"], "kind": full["kind"], "sample": full["sample"],
            "compress_MBps_per_core_loaded": full["compress_MBps_per_core"], "decompress_MBps_per_core_loaded": full["decompress_MBps_per_core"],
            "one_core": {"value": one["value"], "compress_MBps": one["compress_MBps_per_core"], "decompress_MBps": one["decompress_MBps_per_core"]},
            "ratio": full["ratio"], "host": full["host"]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    vals = []
    meas = None
    for i in range(args.warmup + args.steps):
        meas = cpu_measure("lznt1", bytes_per_proc=16 * BLOCK, reps=1)
        if i >= args.warmup:
            vals.append(meas["value"])
    v = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * meas["wall_s"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "LZNT1 encode+decode, 1 GiB synthetic text stream as 4 KiB chunks (BASELINE.json configs[1]); CPU arm: every step times a "
                               "bounded SAMPLE of that stream (16 MiB per host thread, same generator and seeds, one lznt1_compress + lznt1_decompress call per MiB)",
                   "host": meas["host"]},
        "cpu_baseline": {"value": v, "unit": "MB/s", "cores": meas["cores"], "kind": meas["kind"], "sample": meas["sample"]},
        "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
AFFINITY0 = os.sched_getaffinity(0)


def _cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if part:
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(torch, index):
    """Pin this rank (and so its pinned host buffers, by first touch) to the NUMA node the GPU hangs off.  The e2e path moves
    3.4 GB per step per rank through host memory; with 8 ranks on a two-socket box unbound buffers cross the socket link."""
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read())
        if node < 0:
            return {"node": None, "why": "numa_node = -1 (single node or virtualised)"}
        cpus = _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read()) & AFFINITY0
        if not cpus:
            return {"node": node, "why": "no allowed cpu on that node"}
        os.sched_setaffinity(0, cpus)
        try:                                                 # memory policy: prefer the node as well (libnuma, if the box has it)
            numa = ctypes.CDLL("libnuma.so.1")
            if numa.numa_available() >= 0:
                numa.numa_set_preferred(node)
        except OSError:
            pass
        return {"node": node, "cpus": len(cpus)}
    except Exception as e:                                   # no sysfs entry: stay unbound
        return {"node": None, "why": "%s: %s" % (type(e).__name__, e)}

def measure_batch_format(ms, ctx, torch, fmt, each, count, steps, dev, peak, rank, dist=None, world=1):
    """Encode+decode of `count` independent buffers of `each` bytes (generator text, seeds 42+rank*count+i) through the
    device batch API; returns a dict with MB/s, kernel times and roofline fractions."""
    import numpy as np
    U = each * count
    host = np.empty(U, dtype=np.uint8)
    ms.generate_batch_into(host, each, count, 42 + rank * count, "text")
    d_in = torch.empty(U, dtype=torch.uint8, device=dev)
    step_b = 1 << 28
    for o in range(0, U, step_b):
        d_in[o:o + step_b].copy_(torch.from_numpy(host[o:o + step_b]))
    capf = {"xpress_huff": lambda n: n + 34 + 258 + 258 * (n // 65536), "xpress": lambda n: n + 4 + 4 * (n // 32)}[fmt]
    slot = (capf(each) + 15) // 16 * 16
    d_comp = torch.empty(slot * count, dtype=torch.uint8, device=dev)
    d_out = torch.empty(U, dtype=torch.uint8, device=dev)
    ar = torch.arange(count, dtype=torch.int64, device=dev)
    in_off, c_off = ar * each, ar * slot
    in_len = torch.full((count,), each, dtype=torch.int32, device=dev)
    c_cap = torch.full((count,), slot, dtype=torch.int32, device=dev)
    c_len = torch.zeros(count, dtype=torch.int32, device=dev)
    st1 = torch.zeros(count, dtype=torch.int32, device=dev)
    o_len = torch.zeros(count, dtype=torch.int32, device=dev)
    st2 = torch.zeros(count, dtype=torch.int32, device=dev)
    enc = ms.Batch(count, in_off.data_ptr(), in_len.data_ptr(), c_off.data_ptr(), c_cap.data_ptr(), c_len.data_ptr(), st1.data_ptr())
    dec = ms.Batch(count, c_off.data_ptr(), c_len.data_ptr(), in_off.data_ptr(), in_len.data_ptr(), o_len.data_ptr(), st2.data_ptr())

    # plain Xpress: our encoder also writes the side-band restart index (one entry per 16 KiB of input), which lets the
    # decoder parse every segment on its own lane; without it a stream is one lane (reported as decode_ms_unindexed)
    sizes_all = torch.zeros(world, dtype=torch.int64, device=dev)
    # Both formats: our encoder also leaves a side-band restart index (never part of the stream: the streams are byte for byte those
    # of the plain call and decode with the reference), which lets the decoder start inside a stream -- one lane per segment.
    # Without it a stream is one lane (reported as decode_ms_unindexed: the path a reference-made stream takes).
    indexed = True
    seg_bytes = 16384
    stride = (each + seg_bytes - 1) // seg_bytes
    if fmt == "xpress":
        d_index = ctx.xpress_index_tensor(count, stride)
    else:
        d_index, d_ioff, d_nseg = ctx.xhuff_index_alloc(torch.full((count,), each, dtype=torch.int64))

    def step(ev=None):
        if ev:
            ev[0].record()
        if fmt == "xpress":
            ctx.xpress_compress_indexed(d_in, d_comp, enc, seg_bytes, d_index, stride)
        else:
            ctx.xhuff_compress_indexed(d_in, d_comp, enc, d_index, d_ioff, d_nseg)
        if dist is not None:                      # the sharded job's one exchange: every rank learns every rank's compressed size
            ctx.sizes_allgather(c_len.sum(dtype=torch.int64).reshape(1), sizes_all)
        if ev:
            ev[1].record()
        if fmt == "xpress":
            ctx.xpress_decompress_indexed(d_comp, d_out, dec, d_index, stride)
        else:
            ctx.xhuff_decompress_indexed(d_comp, d_out, dec, d_index, d_ioff, d_nseg)
        if ev:
            ev[2].record()

    step()
    torch.cuda.synchronize()
    assert int(st1.abs().sum().item()) == 0 and int(st2.abs().sum().item()) == 0, "%s: kernel reported an error" % fmt
    assert torch.equal(d_out, d_in), "%s round trip mismatch" % fmt
    C = int(c_len.to(torch.int64).sum().item())
    l0 = ctx.launches()
    ctx.profile(True)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    torch.cuda.synchronize()
    for k in range(steps):
        step(evs[k])
    torch.cuda.synchronize()
    prof = ctx.profile_report()
    ctx.profile(False)
    enc_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    dec_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    if dist is not None:                                      # sharded: every rank its own `count` buffers; time = max over ranks
        t = torch.tensor([enc_ms, dec_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        enc_ms, dec_ms = t.tolist()
    # one roofline figure per direction: algorithmic bytes of the whole batch (U + C) over the time of ALL kernels of that direction
    kern = {name: {"calls_per_step": calls / steps, "ms_per_launch": ms_tot / max(calls, 1)} for name, (calls, ms_tot) in prof.items()}
    roof = {d: {"ms": t_ms, "achieved_GBps": (U + C) / 1e9 / (t_ms / 1e3), "frac": (U + C) / 1e9 / (t_ms / 1e3) / peak, "algorithmic_bytes": U + C}
            for d, t_ms in (("encode", enc_ms), ("decode", dec_ms))}
    extra = {}
    if indexed and world == 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.run_batch(fmt, "decompress", d_comp, d_out, dec)          # warm: the first call sizes the token scratch
        d_out.zero_()
        e0.record()
        ctx.run_batch(fmt, "decompress", d_comp, d_out, dec)
        e1.record()
        torch.cuda.synchronize()
        assert torch.equal(d_out, d_in)
        extra = {"decode_ms_unindexed": e0.elapsed_time(e1),
                 "index": ("side-band restart index written by the encoder, one entry per %d input bytes (%d per buffer)" % (seg_bytes, stride)) if fmt == "xpress"
                          else "side-band restart index written by the encoder, one entry (16 bytes) per 64 symbols"}
    if world == 1:
        # the other encoder of the library (mscomp_gpu_set_encoder FAST: bulk-synchronous matcher, valid streams, not the reference's bytes)
        ctx.set_encoder("fast")
        step()
        torch.cuda.synchronize()
        assert int(st1.abs().sum().item()) == 0 and int(st2.abs().sum().item()) == 0 and torch.equal(d_out, d_in), "%s fast encoder round trip" % fmt
        Cf = int(c_len.to(torch.int64).sum().item())
        fe = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(min(steps, 3))]
        for e in fe:
            step(e)
        torch.cuda.synchronize()
        ctx.set_encoder("reference")
        f_enc, f_dec = statistics.mean(e[0].elapsed_time(e[1]) for e in fe), statistics.mean(e[1].elapsed_time(e[2]) for e in fe)
        extra["fast_encoder"] = {"what": "mscomp_gpu_set_encoder(MSCOMP_GPU_ENCODER_FAST): valid streams, not byte-identical to the reference",
                                 "encode_ms": f_enc, "decode_ms": f_dec, "compressed_ratio": Cf / U,
                                 "encode_decode_MBps": U / MB / ((f_enc + f_dec) / 1e3)}
    del d_in, d_comp, d_out
    torch.cuda.empty_cache()
    return {"workload": "%d independent buffers x %d bytes (text) per GPU, device batch API, %d GPU(s)" % (count, each, world),
            "encoder": "reference parse (default): streams byte-identical to the reference's %s_compress" % fmt,
            "uncompressed_bytes_per_gpu": U, "compressed_ratio": C / U,
            **extra, "encode_ms": enc_ms, "decode_ms": dec_ms, "encode_MBps": world * U / MB / (enc_ms / 1e3), "decode_MBps": world * U / MB / (dec_ms / 1e3),
            "encode_decode_MBps": world * U / MB / ((enc_ms + dec_ms) / 1e3), "gpu_launches": ctx.launches() - l0, "roofline": roof, "kernels": kern}


def measure_mixed(ms, ctx, torch, dist, dev, rank, world, count, each, steps):
    """BASELINE.json configs[4]: `count` independent buffers of `each` bytes per GPU, buffer i (global index) in format
    [LZNT1, Xpress, Xpress-Huffman][i % 3], generator text with seed 42 + i.  Buffers are sharded contiguously by rank; the only
    cross-GPU traffic is one all_gather of the per-rank compressed sizes.  Encode and decode of all three groups are timed
    together with CUDA events; every round trip is verified."""
    import numpy as np
    U = count * each
    host = np.empty(U, dtype=np.uint8)
    ms.generate_batch_into(host, each, count, 42 + rank * count, "text")
    g = [[j for j in range(count) if (rank * count + j) % 3 == f] for f in range(3)]
    d_all = torch.empty(U, dtype=torch.uint8, device=dev)
    for o in range(0, U, 1 << 28):
        d_all[o:o + (1 << 28)].copy_(torch.from_numpy(host[o:o + (1 << 28)]))
    view = d_all.view(count, each)
    # LZNT1 group: its buffers back to back are one stream of independent 4 KiB chunks (each is a multiple of 4096)
    d_l = view[torch.tensor(g[0], device=dev)].contiguous().view(-1)
    nl = d_l.numel()
    # Xpress group: whole buffers with the restart index; Xpress-Huffman group: 64 KiB units (the format's chunk)
    def batch_for(idx, unit):
        per = each // unit
        offs = torch.tensor([j * each + k * unit for j in idx for k in range(per)], dtype=torch.int64, device=dev)
        n = offs.numel()
        capf = (lambda m: m + 4 + 4 * (m // 32)) if unit == each else (lambda m: m + 34 + 258 + 258 * (m // 65536))
        slot = (capf(unit) + 15) // 16 * 16
        c_off = torch.arange(n, dtype=torch.int64, device=dev) * slot
        lens = torch.full((n,), unit, dtype=torch.int32, device=dev)
        caps = torch.full((n,), slot, dtype=torch.int32, device=dev)
        c_len = torch.zeros(n, dtype=torch.int32, device=dev)
        st1 = torch.zeros(n, dtype=torch.int32, device=dev)
        o_len = torch.zeros(n, dtype=torch.int32, device=dev)
        st2 = torch.zeros(n, dtype=torch.int32, device=dev)
        d_comp = torch.empty(slot * n, dtype=torch.uint8, device=dev)
        enc = ms.Batch(n, offs.data_ptr(), lens.data_ptr(), c_off.data_ptr(), caps.data_ptr(), c_len.data_ptr(), st1.data_ptr())
        dec = ms.Batch(n, c_off.data_ptr(), c_len.data_ptr(), offs.data_ptr(), lens.data_ptr(), o_len.data_ptr(), st2.data_ptr())
        return {"n": n, "enc": enc, "dec": dec, "comp": d_comp, "c_len": c_len, "st": (st1, st2), "keep": (offs, c_off, lens, caps, o_len)}
    bx, bh = batch_for(g[1], each), batch_for(g[2], 65536)
    seg_bytes, stride = 16384, (each + 16383) // 16384
    d_index = ctx.xpress_index_tensor(bx["n"], stride)
    h_index, h_ioff, h_nseg = ctx.xhuff_index_alloc(torch.full((bh["n"],), 65536, dtype=torch.int64))
    d_out = torch.zeros(U, dtype=torch.uint8, device=dev)
    state = {}

    # The three groups are independent: each runs on its own stream (own context), so the latency-bound lane kernels of one
    # format overlap with the kernels of the others.  Events on the main stream bracket all of them.
    main = torch.cuda.current_stream()
    side = [torch.cuda.Stream() for _ in range(3)]
    cx = []
    for st_ in side:
        with torch.cuda.stream(st_):
            cx.append(ms.GpuContext())
    fork, join = torch.cuda.Event(), [torch.cuda.Event() for _ in range(3)]

    def phase(fns):
        fork.record(main)
        for i, fn in enumerate(fns):
            side[i].wait_event(fork)
            with torch.cuda.stream(side[i]):
                fn()
                join[i].record(side[i])
        for i in range(3):
            main.wait_event(join[i])

    def enc_l():
        state["l"] = cx[0].lznt1_compress(d_l)

    def dec_l():
        d_lc, d_loff, _ = state["l"]
        state["lo"] = cx[0].lznt1_decompress(d_lc, d_loff, nl // 4096)

    def step(ev=None):
        if ev:
            ev[0].record()
        phase([enc_l, lambda: cx[1].xpress_compress_indexed(d_all, bx["comp"], bx["enc"], seg_bytes, d_index, stride),
               lambda: cx[2].xhuff_compress_indexed(d_all, bh["comp"], bh["enc"], h_index, h_ioff, h_nseg)])
        if ev:
            ev[1].record()
        phase([dec_l, lambda: cx[1].xpress_decompress_indexed(bx["comp"], d_out, bx["dec"], d_index, stride),
               lambda: cx[2].xhuff_decompress_indexed(bh["comp"], d_out, bh["dec"], h_index, h_ioff, h_nseg)])
        if ev:
            ev[2].record()

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    for b in (bx, bh):
        assert int(b["st"][0].abs().sum().item()) == 0 and int(b["st"][1].abs().sum().item()) == 0, "mixed batch: kernel reported an error"
    ov = d_out.view(count, each)
    for f in (1, 2):
        ix = torch.tensor(g[f], device=dev)
        assert torch.equal(ov[ix], view[ix]), "mixed batch: round trip mismatch (format %d)" % f
    assert torch.equal(state["lo"][0][:nl], d_l), "mixed batch: LZNT1 round trip mismatch"
    C = [int(ms.parse_result(state["l"][2])["bytes"]), int(bx["c_len"].to(torch.int64).sum().item()), int(bh["c_len"].to(torch.int64).sum().item())]
    l0 = sum(c.launches() for c in cx)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    for k in range(steps):
        step(evs[k])
        if dist is not None:                      # the one collective of the sharded job: everybody learns every rank's output sizes
            sizes = torch.tensor(C, dtype=torch.int64, device=dev)
            gathered = [torch.empty_like(sizes) for _ in range(world)]
            dist.all_gather(gathered, sizes)
    torch.cuda.synchronize()
    t = torch.tensor([statistics.mean(e[0].elapsed_time(e[1]) for e in evs), statistics.mean(e[1].elapsed_time(e[2]) for e in evs)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    enc_ms, dec_ms = t.tolist()
    launches = sum(c.launches() for c in cx) - l0
    del d_all, d_l, d_out, bx, bh, state
    torch.cuda.empty_cache()
    return {"config": "BASELINE.json configs[4]: %d independent %d-byte buffers per GPU, format [LZNT1, Xpress, Xpress-Huffman][i %% 3], sharded contiguously over %d GPU(s); "
                      "Xpress-Huffman buffers as 64 KiB units, Xpress and Xpress-Huffman with the side-band restart index their encoders write" % (count, each, world),
            "uncompressed_bytes_per_gpu": U, "compressed_ratio": sum(C) / U, "encode_ms": enc_ms, "decode_ms": dec_ms,
            "encode_decode_MBps": world * U / MB / ((enc_ms + dec_ms) / 1e3), "encode_MBps": world * U / MB / (enc_ms / 1e3),
            "decode_MBps": world * U / MB / (dec_ms / 1e3), "gpu_launches": launches,
            "collective": "all_gather of 3 x int64 per rank per step" if world > 1 else "none (N=1)",
            "streams": "one CUDA stream per format group (3), bracketed by events on the main stream"}


def measure_oneshot(ms, mib=8, reps=3):
    """The drop-in calls on ONE buffer in ordinary (pageable) host memory -- what a user of the reference's library hits first:
    <fmt>_compress / <fmt>_decompress of one `mib` MiB text buffer, best of `reps`, only the C call inside the timer (buffers are
    allocated and touched before).  The default encoders are byte-identical to the reference's, so the decoders are timed on
    reference-made streams."""
    import numpy as np
    n = mib << 20
    plain = np.frombuffer(ms.generate("text", n, 42), dtype=np.uint8).copy()
    L = ms.lib()
    out = {"what": "one %d MiB text buffer per call, pageable host memory, C ABI one-shot calls, best of %d; MB/s of uncompressed bytes" % (mib, reps)}
    for fmt, cap in (("lznt1", n + 3 + 2 * ((n + 4095) // 4096)), ("xpress", n + 4 + 4 * (n // 32)), ("xpress_huff", n + 34 + 258 + 258 * (n // 65536))):
        comp = np.zeros(cap + 64, dtype=np.uint8)
        back = np.zeros(n + 64, dtype=np.uint8)
        tc, td = [], []
        clen = 0
        for _ in range(reps):
            m = ctypes.c_size_t(cap)
            t0 = time.perf_counter()
            st = getattr(L, fmt + "_compress")(ctypes.c_void_p(plain.ctypes.data), ctypes.c_size_t(n), ctypes.c_void_p(comp.ctypes.data), ctypes.byref(m))
            t1 = time.perf_counter()
            assert st == 0, (fmt, st)
            clen = m.value
            k = ctypes.c_size_t(n)
            t2 = time.perf_counter()
            st = getattr(L, fmt + "_decompress")(ctypes.c_void_p(comp.ctypes.data), ctypes.c_size_t(clen), ctypes.c_void_p(back.ctypes.data), ctypes.byref(k))
            t3 = time.perf_counter()
            assert st == 0 and k.value == n, (fmt, st, k.value)
            tc.append(t1 - t0); td.append(t3 - t2)
        assert bytes(back[:n]) == bytes(plain), fmt + " one-shot round trip"
        out[fmt] = {"compress_ms": 1e3 * min(tc), "decompress_ms": 1e3 * min(td), "compress_MBps": n / MB / min(tc), "decompress_MBps": n / MB / min(td),
                    "compressed_ratio": clen / n}
    return out


def run_gpu(args):
    import numpy as np
    import torch
    import mscompress_b200 as ms

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)              # before any pinned allocation: first touch lands on the GPU's node
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    nblocks = args.mib
    U = nblocks * BLOCK
    host = torch.empty(U, dtype=torch.uint8).pin_memory()
    ms.generate_batch_into(host.numpy(), BLOCK, nblocks, 42 + rank * nblocks, "text")
    d_in = host.to(dev, non_blocking=True)
    ctx = ms.GpuContext(local)
    if dist is not None:
        # the data-path exchange goes through the LIBRARY's multi-GPU entry (C ABI: mscomp_gpu_lznt1_compress_sharded = single-GPU
        # pipeline + ncclAllGather on the compute stream, its own communicator); torch.distributed only carries the NCCL id and the timers
        from importlib import import_module
        ctx.comm_init(rank, world, import_module("mscompress_b200.sharding").torch_broadcaster(dist, dev))
    nchunks = U // 4096
    d_comp = torch.empty(U + 2 * nchunks + 64, dtype=torch.uint8, device=dev)
    d_off = torch.empty(nchunks + 1, dtype=torch.int64, device=dev)
    d_out = torch.empty(U, dtype=torch.uint8, device=dev)
    d_res_c = torch.zeros(ms.RESULT_BYTES, dtype=torch.uint8, device=dev)
    d_res_d = torch.zeros(ms.RESULT_BYTES, dtype=torch.uint8, device=dev)
    sizes_all = torch.zeros(world, dtype=torch.int64, device=dev)

    def step(events=None):
        if events:
            events[0].record()
        if dist is not None:   # the single collective -- all-gather of the per-rank compressed sizes, inside the library call
            ctx.lznt1_compress_sharded(d_in, d_comp, d_off, d_res_c, sizes_all)
        else:
            ctx.lznt1_compress(d_in, d_comp, d_off, d_res_c)
        if events:
            events[1].record()
        # the compressed stream stays resident; its length lives in d_off[nchunks] / d_res_c (the decoder takes the buffer bound)
        ctx.lznt1_decompress(d_comp, d_off, nchunks, d_out, d_res_d)
        if events:
            events[2].record()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    C = ms.parse_result(d_res_c)["bytes"]
    rd = ms.parse_result(d_res_d)
    assert rd["status"] == 0 and rd["bytes"] == U and torch.equal(d_out, d_in), "round trip mismatch"

    launches0 = ctx.launches()
    ctx.profile(True)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    with Clocks(local) as clk:
        start.record()
        for k in range(args.steps):
            step(ev[k])
        end.record()
        torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    total_ms = start.elapsed_time(end)
    prof = ctx.profile_report()
    ctx.profile(False)
    launches = ctx.launches() - launches0
    enc_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
    dec_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
    t = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    value = world * U / MB / (ms_per_step / 1e3)

    # ---- e2e through the C ABI with pinned host buffers ----
    h_comp = torch.empty(U + 2 * nchunks + 64, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(U, dtype=torch.uint8).pin_memory()
    L = ms.lib()
    e2e_steps = max(1, args.steps)
    e2e_t = []
    csz = 0
    for k in range(1 + e2e_steps):
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        n = ctypes.c_size_t(h_comp.numel())
        st = L.lznt1_compress(host.data_ptr(), U, h_comp.data_ptr(), ctypes.byref(n))
        m = ctypes.c_size_t(U)
        st2 = L.lznt1_decompress(h_comp.data_ptr(), n.value, h_out.data_ptr(), ctypes.byref(m))
        dt = time.perf_counter() - t0
        assert st == 0 and st2 == 0 and m.value == U, (st, st2)
        csz = n.value
        if k > 0:
            e2e_t.append(dt)
    assert torch.equal(h_out, host) and csz == C
    te = torch.tensor([statistics.mean(e2e_t)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    peak, peak_src = peaks()
    formats = {}
    if args.formats != "none":
        del d_comp, d_out, h_comp, h_out
        torch.cuda.empty_cache()
        gib = args.xhuff_gib
        fsteps = max(2, min(args.steps, 5))
        formats["xpress_huff"] = measure_batch_format(ms, ctx, torch, "xpress_huff", 65536, gib * 16384, fsteps, dev, peak, rank, dist, world)
        formats["xpress_huff"]["config"] = ("BASELINE.json configs[3]: %d GiB synthetic text per GPU as independent 64 KiB buffers, "
                                            "chunk-sharded over %d GPU(s) (weak scaling)" % (gib, world))
        formats["xpress"] = measure_batch_format(ms, ctx, torch, "xpress", BLOCK, args.mib, fsteps, dev, peak, rank, dist, world)
        formats["xpress"]["config"] = ("BASELINE.json configs[2]: %d MiB synthetic text per GPU as %d independent 1 MiB buffers (one stream each), "
                                       "%d GPU(s)" % (args.mib, args.mib, world))

    mixed = None
    if args.formats != "none" and hasattr(ms.lib(), "mscomp_gpu_xpress_compress_batch_indexed"):
        try:
            del d_comp, d_out, h_comp, h_out
        except NameError:
            pass
        torch.cuda.empty_cache()
        mixed = measure_mixed(ms, ctx, torch, dist, dev, rank, world, 1024, BLOCK, 3)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    steps = args.steps
    kern = {}
    alg = {"lznt1_encode_kernel": U + C, "lznt1_decode_kernel": C + U, "gather_slots_kernel": 2 * C}
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    for name, (calls, ms_tot) in prof.items():
        per = ms_tot / max(calls, 1)
        e = {"calls_per_step": calls / steps, "ms_per_launch": per}
        if name in alg and per > 0:
            e["achieved_GBps"] = alg[name] / 1e9 / (per / 1e3)
            e["frac"] = e["achieved_GBps"] / peak
        kern[name] = e
    dec_name = "lznt1_decode_kernel"
    dom = max((k for k in kern if k in alg), key=lambda k: kern[k]["ms_per_launch"] * kern[k]["calls_per_step"])
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kern[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac"],
                "traffic": traffic.get(dom), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "note": "achieved = (uncompressed + compressed bytes) / CUDA-event kernel time (SURVEY.md 8d)"}
    cpu = cpu_fmt = None
    if world == 1:
        os.sched_setaffinity(0, AFFINITY0)                   # the CPU arm gets every host thread back (the GPU arm was bound to one NUMA node)
        cpu = cpu_baseline_block("lznt1", 8 * BLOCK)
        if formats:
            cpu_fmt = {"xpress": cpu_baseline_block("xpress", 16 * BLOCK), "xpress_huff": cpu_baseline_block("xpress_huff", 8 * BLOCK)}

    line = {
        "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "LZNT1 encode+decode, %d MiB synthetic text stream as 4 KiB chunks per GPU (BASELINE.json configs[1])" % nblocks,
                   "chunks_per_gpu": nchunks, "compressed_ratio": C / U, "l2": "inputs larger than L2 (1 GiB in, %.2f GiB compressed, 126 MB L2); no flush needed" % (C / 2**30),
                   "collective": ("ncclAllGather of the per-rank compressed size (8 B/rank) per step, issued by mscomp_gpu_lznt1_compress_sharded on the "
                                  "compute stream (%d collectives, communicator of %d ranks)" % (ctx.collectives(), ctx.comm_size())) if world > 1 else "none (N=1)"},
        "gpu_launches": launches,
        "encode_MBps": world * U / MB / (enc_ms / 1e3), "decode_MBps": world * U / MB / (dec_ms / 1e3),
        "encode_ms": enc_ms, "decode_ms": dec_ms,
        "roofline": roofline, "kernels": kern,
        "decode_roofline": {"kernel": dec_name, "bound": "hbm", "achieved": kern.get(dec_name, {}).get("achieved_GBps"), "peak": peak, "unit": "GB/s",
                            "frac": kern.get(dec_name, {}).get("frac"), "target_frac": 0.40},
        "e2e": {"value": world * U / MB / e2e_s, "unit": "MB/s", "h2d_bytes_per_step": U + C, "d2h_bytes_per_step": C + U,
                "api": "lznt1_compress + lznt1_decompress (C ABI, pinned host buffers)", "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps,
                "numa": numa},
        "clocks": clk.summary(),
        "formats": formats,
    }
    if mixed:
        line["mixed_batch"] = mixed
    if cpu:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline_detail"] = {k: v for k, v in cpu.items() if k not in line["cpu_baseline"]}
    if cpu_fmt:
        for f, c in cpu_fmt.items():
            line["formats"][f]["cpu_baseline"] = c
            line["formats"][f]["vs_cpu_all_cores"] = line["formats"][f]["encode_decode_MBps"] / c["value"]
    if world == 1:
        one = measure_oneshot(ms)
        ref1 = {"lznt1": (cpu or {}).get("one_core")}
        for f, c in (cpu_fmt or {}).items():
            ref1[f] = c.get("one_core")
        for f, r in ref1.items():
            if r and f in one:
                one[f]["reference_one_core"] = {"compress_MBps": r.get("compress_MBps"), "decompress_MBps": r.get("decompress_MBps")}
        line["oneshot"] = one
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line):
    """The one JSON line of the contract, on the process's ORIGINAL stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print on stdout too (NCCL's version banner at any NCCL_DEBUG level, torch warnings): send file descriptor 1
    # to stderr for the whole run and keep the original for the JSON line, so stdout carries exactly one line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--formats", default="all", choices=["all", "none"], help="also measure Xpress-Huffman (configs[3]) and Xpress (configs[2]) at N=1")
    ap.add_argument("--xhuff-gib", type=int, default=4, help="GiB of uncompressed data for the Xpress-Huffman measurement")
    ap.add_argument("--mib", type=int, default=1024, help="MiB of uncompressed data per GPU (default: the 1 GiB of BASELINE.json configs[1])")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
