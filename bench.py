#!/usr/bin/env python
"""bench.py -- audio-seconds/sec of 128-bin log-mel fbank per B200 (BASELINE.json metric).

One "step" = one pass of the fused hot path over one batch of synthetic clips:
``configs[1]`` of BASELINE.json -- 1024 ESC-50-shaped clips (5 s, 44.1 kHz mono float32)
-> 16 kHz kaldi fbank (hanning, 128 mel, 10 ms hop) -> 512-frame target -> mean/std
normalisation, float32.  Under torchrun every rank processes its own 1024 clips
(weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

Prints ONE JSON line (rank 0).  ``value`` = device-resident throughput (CUDA events, max
over ranks); ``e2e`` = same metric through the public API with HOST buffers (H2D of the
waveforms and D2H of the features inside the timed region); ``roofline`` = algorithmic
bytes / kernel time vs the measured HBM copy peak; ``cpu_baseline`` = the reference CPU
path (torchaudio Resample + kaldi.fbank) timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

CLIP_SAMPLES = 220500          # 5 s @ 44.1 kHz
CLIP_SECONDS = 5.0
OUT_FRAMES = 512
N_MELS = 128
AST_MEAN, AST_STD = -6.6268, 5.0613      # SURVEY.md section 6 / BASELINE.md section 4
METRIC = "audio-sec/sec log-mel fbank"
UNIT = "audio-s/s"
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback


def measured_traffic():
    """DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r02_ws_full_metrics.csv: dram__bytes_read.sum + dram__bytes_write.sum), or None.  ncu cannot run inside a
    timed bench, so this is the one figure of the line that is not measured live."""
    try:
        tot = 0.0
        name = "r02_ws_full_metrics.csv" if os.path.exists(os.path.join(ROOT, "profiles", "r02_ws_full_metrics.csv")) else "r01c_ws_full_metrics.csv"
        for line in open(os.path.join(ROOT, "profiles", name)):
            f = line.strip().split(",")
            if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
        return tot or None
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


# --------------------------------------------------------------------------------------
# CPU reference arm: the reference's own path (torchaudio on host cores)
# --------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One DataLoader-style worker: n clips through Resample -> kaldi.fbank -> pad -> norm."""
    n, seed, kind = args
    import numpy as np
    import torch
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(seed)
    def fresh(k):                                     # k NEW synthetic clips (generated outside the timed spans)
        return [torch.rand(1, CLIP_SAMPLES, generator=g) * 2 - 1 for _ in range(k)]
    clips = fresh(1)
    if kind == "reference":
        import torchaudio.compliance.kaldi as kaldi
        import torchaudio.transforms as T
        rs = T.Resample(44100, 16000)

        def one(w):
            f = kaldi.fbank(rs(w), htk_compat=True, sample_frequency=16000, use_energy=False,
                            window_type="hanning", num_mel_bins=128, dither=0.0, frame_shift=10)
            f = torch.nn.functional.pad(f, (0, 0, 0, OUT_FRAMES - f.shape[0]))
            return (f - AST_MEAN) / (2 * AST_STD)
    else:
        from oracle import fbank_oracle as O

        def one(w):
            return O.ast_frontend(w[0].numpy(), 44100, target_frames=OUT_FRAMES, mean=AST_MEAN, std=AST_STD)[0]
    one(clips[0])                                     # warm the worker (imports, filter tables)
    busy, done = 0.0, 0
    while done < n:                                   # every clip is distinct; only the transform is timed
        clips = fresh(min(8, n - done))
        t0 = time.perf_counter()
        for w in clips:
            one(w)
        busy += time.perf_counter() - t0
        done += len(clips)
    return busy


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _cpu_single(args):
    """BASELINE.md section 4 modes 1 and 2: ONE process with `threads` intra-op threads over n clips (run in a child so
    that the thread setting does not leak into the parent)."""
    n, threads, actual = args
    import torch
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    clips = [torch.rand(1, CLIP_SAMPLES, generator=g) * 2 - 1 for _ in range(8)]
    import torchaudio.compliance.kaldi as kaldi
    import torchaudio.transforms as T
    if actual:          # the reference-actual recipe (src/datasets/preprocessing.py:988-998, 1013-1039), restated with torchaudio
        mel = T.MelSpectrogram(sample_rate=44100, n_fft=1024, win_length=400, hop_length=160, n_mels=128, power=2.0)
        db = T.AmplitudeToDB(top_db=80)

        def one(w):
            x = db(mel(w))
            return (x - x.mean()) / x.std() * 0.5
    else:
        rs = T.Resample(44100, 16000)

        def one(w):
            f = kaldi.fbank(rs(w), htk_compat=True, sample_frequency=16000, use_energy=False,
                            window_type="hanning", num_mel_bins=128, dither=0.0, frame_shift=10)
            f = torch.nn.functional.pad(f, (0, 0, 0, OUT_FRAMES - f.shape[0]))
            return (f - AST_MEAN) / (2 * AST_STD)
    one(clips[0])
    t0 = time.perf_counter()
    for i in range(n):
        one(clips[i % 8])
    return n * CLIP_SECONDS / (time.perf_counter() - t0)


def cpu_modes(n=40):
    """audio-s/s of BASELINE.md section 4 modes 1 (1 thread) and 2 (all cores, intra-op) and of the reference-actual recipe
    (all cores), each over n clips in a forked child; call BEFORE CUDA is initialised."""
    import multiprocessing as mp
    import torch  # noqa: F401
    try:
        import torchaudio.compliance.kaldi  # noqa: F401
    except Exception:
        return None
    ctx = mp.get_context("fork")
    cores = os.cpu_count() or 1
    out = {}
    with ctx.Pool(1) as pool:
        out["mode1_1proc_1thread"] = pool.apply(_cpu_single, ((n, 1, False),))
    with ctx.Pool(1) as pool:
        out["mode2_1proc_allthreads"] = pool.apply(_cpu_single, ((n, cores, False),))
    with ctx.Pool(1) as pool:
        out["reference_actual_melspec_db_allthreads"] = pool.apply(_cpu_single, ((max(8, n // 2), cores, True),))
    return out


def cpu_reference_throughput(clips_per_worker=300, workers=None):
    """audio-s/s of the CPU path with `workers` single-thread processes (the reference's
    DataLoader(num_workers) pattern, configs/base_training.yaml:104; workers are forked
    like DataLoader workers, so call this BEFORE CUDA is initialised).  Returns dict."""
    import multiprocessing as mp
    import torch  # noqa: F401  (imported in the parent so forked workers inherit it)
    try:
        import torchaudio.compliance.kaldi  # noqa: F401
        import torchaudio.transforms  # noqa: F401
        kind = "reference"
    except Exception:
        kind = "port"
    workers = workers or os.cpu_count() or 1
    per = max(1, clips_per_worker)
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        t_spawn = time.perf_counter() - t0
        t1 = time.perf_counter()
        busy = pool.map(_cpu_worker, [(per, 1234 + i, kind) for i in range(workers)])
        wall = time.perf_counter() - t1
    # throughput over the slowest worker's compute time (excludes interpreter spawn + import)
    t = max(busy)
    n = per * workers
    return dict(value=n * CLIP_SECONDS / t, unit=UNIT, cores=workers, kind=kind,
                sample=f"{n} clips x 5 s @44.1 kHz ({per}/worker, {workers} single-thread worker processes; "
                       f"slowest worker {t:.2f} s, pool wall {wall:.2f} s, spawn {t_spawn:.2f} s); "
                       f"torchaudio Resample(44100,16000) + kaldi.fbank + pad 512 + (x-mean)/(2 std)"
                       if kind == "reference" else
                       f"{n} clips x 5 s via the numpy oracle port ({workers} workers)",
                seconds=t, clips=n)


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                    reasons=sorted(reasons))


# --------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    # bounded sample: 200 clips per worker per step (~1-2 s of host work per step)
    workers = os.cpu_count() or 1
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_reference_throughput(200, workers)
        if i >= args.warmup:
            times.append(res["seconds"])
    t = sum(times) / len(times)
    value = res["clips"] * CLIP_SECONDS / t
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=f"AST frontend, {res['clips']} ESC-50 clips (5 s @44.1 kHz) per step -> 16 kHz kaldi "
                                     f"fbank 128 mel, 512-frame target, mean/std normalisation (bounded sample of the "
                                     f"batch-1024 config)", l2="n/a (CPU)"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=res["cores"], kind=res["kind"], sample=res["sample"]),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def run_b200(args, rank, local_rank, world):
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_reference_throughput(300)           # before CUDA init: workers are forked
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu["mode"] = "3: os.cpu_count() single-thread worker processes (BASELINE.md section 4)"
        cpu["cpu_model"] = cpu_model()
        cpu["other_modes_audio_s_per_s"] = cpu_modes(40)
    import torch
    import torch.distributed as dist
    import dl_sound_classification_b200 as b2

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    fe = b2.FbankFrontend(orig_rates=(44100,), device=dev, **b2.AST_FBANK_KWARGS)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    wav = torch.rand((B, CLIP_SAMPLES), generator=gen, device=dev, dtype=torch.float32) * 2 - 1
    out = torch.empty((B, OUT_FRAMES, N_MELS), device=dev, dtype=torch.float32)
    mean = torch.tensor([AST_MEAN], device=dev)
    std = torch.tensor([AST_STD], device=dev)

    def step():
        fe(wav, out_frames=OUT_FRAMES, mean=mean, std=std, out=out, return_n_frames=False)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.15)
    b2.launch_count(reset=True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    barrier()
    launches = b2.launch_count()
    total_ms = evs[0].elapsed_time(evs[-1])
    kernel_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    if sampler:
        time.sleep(0.1)
        clocks = sampler.stop()
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * CLIP_SECONDS * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers in, host buffers out, through the public API ----------------
    h_wav = torch.empty((B, CLIP_SAMPLES), dtype=torch.float32).pin_memory()
    h_wav.copy_(wav)
    h_out = torch.empty((B, OUT_FRAMES, N_MELS), dtype=torch.float32).pin_memory()

    def e2e_step():      # public host-in / host-out call: chunked H2D -> kernel -> D2H over three streams
        fe.process_host(h_wav, OUT_FRAMES, h_out=h_out, chunk_clips=args.e2e_chunk, mean=mean, std=std)

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / e2e_steps
    e2e_value = world * B * CLIP_SECONDS / (e2e_ms * 1e-3)
    checksum = float(h_out[0, :498].double().sum())

    # the same call fed with 16-bit PCM host buffers (what an ESC-50 WAV holds): half the host -> device bytes, widened on
    # the device bit-identically to torchaudio.load's float32 (tests/test_gpu_clip_norm.py)
    h_pcm = (h_wav * 32767.0).round().to(torch.int16).pin_memory()

    def e2e_pcm_step():
        fe.process_host(h_pcm, OUT_FRAMES, h_out=h_out, chunk_clips=args.e2e_chunk, mean=mean, std=std)

    for _ in range(2):
        e2e_pcm_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_pcm_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pcm_ms = float(t.item()) / e2e_steps
    del h_pcm, h_wav

    extra = run_extras(args, torch, dist, b2, dev, rank, world, barrier) if not args.no_extra else None

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg_bytes = B * (CLIP_SAMPLES * 4 + OUT_FRAMES * N_MELS * 4)
        k_ms = sum(kernel_ms) / len(kernel_ms)
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="f32", data="synthetic",
            config=dict(workload=f"AST frontend batch {B} ESC-50 clips (5 s @44.1 kHz mono f32) per GPU -> polyphase "
                                 f"resample to 16 kHz -> kaldi fbank (hanning, 128 mel, 25/10 ms, 512-pt FFT) -> pad to "
                                 f"512 frames -> (x-mean)/(2 std); BASELINE.json configs[1]",
                        batch_per_gpu=B, clip_samples=CLIP_SAMPLES, out_frames=OUT_FRAMES, n_mels=N_MELS,
                        l2=f"inputs {B * CLIP_SAMPLES * 4 / 1e6:.0f} MB per step exceed the 126 MB L2 (no flush needed)",
                        kernel=os.environ.get("B200FBANK_KERNEL", "auto")),
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                          traffic=measured_traffic() if B == 1024 else None, peak_source=peak_src, algorithmic_bytes_per_launch=alg_bytes,
                          kernel_ms=k_ms),
            cpu_baseline=cpu,
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=B * CLIP_SAMPLES * 4,
                     d2h_bytes_per_step=B * OUT_FRAMES * N_MELS * 4, steps=e2e_steps, checksum=checksum,
                     ms_per_step=e2e_ms, h2d_gbs_per_rank=B * CLIP_SAMPLES * 4 / (e2e_ms * 1e-3) / 1e9,
                     d2h_gbs_per_rank=B * OUT_FRAMES * N_MELS * 4 / (e2e_ms * 1e-3) / 1e9,
                     host_input="float32 pinned (B, 220500)"),
            e2e_pcm16=dict(value=world * B * CLIP_SECONDS / (pcm_ms * 1e-3), unit=UNIT, ms_per_step=pcm_ms,
                           h2d_bytes_per_step=B * CLIP_SAMPLES * 2, d2h_bytes_per_step=B * OUT_FRAMES * N_MELS * 4,
                           h2d_gbs_per_rank=B * CLIP_SAMPLES * 2 / (pcm_ms * 1e-3) / 1e9,
                           host_input="int16 PCM pinned (B, 220500), widened on the device (b200fbank_pcm16_to_float)"),
            extra=extra,
            gpu_launches=launches, clocks=clocks)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(args, torch, dist, b2, dev, rank, world, barrier):
    """Short runs of BASELINE.json configs[2], [3], [4] at this world size, carried in the headline line's `extra`
    object so that the driver's BENCH / SCALE records hold them at every N (all ranks take part; times are the max over
    ranks).  us8k and sweep are weak scaling (every rank its own batch); stats shards 100 000 clips over the ranks and has
    the NCCL all-reduce INSIDE the timed span."""
    import random as _random
    from dl_sound_classification_b200 import stats as ST
    peak = measured_peaks()[0]

    def timed(fn, steps, warmup=3):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    out = {}
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    mean, std = torch.tensor([AST_MEAN], device=dev), torch.tensor([AST_STD], device=dev)
    # ---- configs[2]: US8K-shaped ragged batch ------------------------------------------------
    B8, table = 4096, (22050, 44100, 48000)
    g = torch.Generator().manual_seed(31 + rank)
    rid = torch.randint(0, 3, (B8,), generator=g)
    lens = ((1.0 + 3.0 * torch.rand(B8, generator=g)) * torch.tensor(table)[rid]).long()
    offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).to(dev)
    flat = torch.rand(int(offsets[-1]), generator=gen, device=dev) * 2 - 1
    fe8 = b2.FbankFrontend(orig_rates=table, device=dev, **b2.AST_FBANK_KWARGS)
    _random.seed(77)
    masks = b2.specaugment.draw_masks(B8, 1024, 128, 192, 48).to(dev)
    rid_d = rid.int().to(dev)
    out8 = torch.empty((B8, 1024, N_MELS), device=dev)
    nfr = fe8(flat, 1024, offsets=offsets, rate_ids=rid_d, masks=masks, mean=mean, std=std, out=out8)[1]
    ms = timed(lambda: fe8(flat, 1024, offsets=offsets, rate_ids=rid_d, masks=masks, mean=mean, std=std, out=out8,
                           return_n_frames=False), 20)
    secs = float((lens.double() / torch.tensor(table, dtype=torch.float64)[rid]).sum())
    in_bytes = int(lens.sum()) * 4
    real_rows = int(nfr.sum())
    out["us8k"] = dict(workload="configs[2]: 4096 ragged clips per GPU (1-4 s @22.05/44.1/48 kHz) -> 1024 frames, SpecAugment masks, mean/std",
                       ms_per_step=ms, value=world * secs / (ms * 1e-3), unit=UNIT, audio_seconds_per_step=secs,
                       roofline_frac_padded_rows=(in_bytes + B8 * 1024 * N_MELS * 4) / (ms * 1e-3) / 1e9 / peak,
                       roofline_frac_real_rows=(in_bytes + real_rows * N_MELS * 4) / (ms * 1e-3) / 1e9 / peak,
                       real_rows_fraction=real_rows / (B8 * 1024.0))
    del flat, out8, fe8
    # ---- configs[3]: dataset statistics, 100 000 clips sharded over the ranks + all-reduce -----
    total = args.clips
    lo, hi = ST.shard_bounds(total, rank, world)
    chunk = 4096
    fe = b2.FbankFrontend(orig_rates=(44100,), device=dev, **b2.AST_FBANK_KWARGS)
    wav = torch.rand((chunk, CLIP_SAMPLES), generator=gen, device=dev) * 2 - 1     # one synthetic chunk, re-used (untimed set-up)
    ds = b2.DatasetStats(fe, OUT_FRAMES)
    ds.update(wav[:64]); ds.all_reduce()                                           # warm-up (NCCL communicator included)
    ds = b2.DatasetStats(fe, OUT_FRAMES)
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for c0 in range(lo, hi, chunk):
        ds.update(wav[:min(chunk, hi - c0)])
    e1.record()
    ds.all_reduce()
    e2.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e2), e1.elapsed_time(e2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = ds.finalize()
    out["stats"] = dict(workload=f"configs[3]: per-bin sum / sum of squares over {total} synthetic ESC-50 clips sharded over {world} GPU(s), "
                                 f"NCCL all-reduce of 257 doubles inside the timed span",
                        ms_total=float(t[0]), ms_allreduce=float(t[1]), value=total * CLIP_SECONDS / (float(t[0]) * 1e-3), unit=UNIT,
                        frames=st.frames, roofline_frac=total * CLIP_SAMPLES * 4 / world / (float(t[0]) * 1e-3) / 1e9 / peak)
    # ---- configs[4]: batch sweep ------------------------------------------------------------
    sweep = []
    out_b = torch.empty((chunk, OUT_FRAMES, N_MELS), device=dev)
    for B in (64, 256, 1024, 4096, 16384, 65536):
        nb = min(B, chunk)
        reps = B // nb

        def step():
            for _ in range(reps):                                  # > 4096 clips: streamed in 4096-clip launches
                fe(wav[:nb], OUT_FRAMES, mean=mean, std=std, out=out_b[:nb], return_n_frames=False)
        ms = timed(step, max(3, 40 // reps))
        sweep.append(dict(batch_per_gpu=B, ms_per_step=ms, value=world * B * CLIP_SECONDS / (ms * 1e-3),
                          roofline_frac=B * (CLIP_SAMPLES * 4 + OUT_FRAMES * N_MELS * 4) / (ms * 1e-3) / 1e9 / peak))
    out["sweep"] = dict(workload="configs[4]: batch 64 .. 65536 clips per GPU (beyond 4096: streamed 4096-clip launches)", unit=UNIT, points=sweep)
    # ---- per-clip statistics instead of dataset statistics (the reference's own normalisation) ----
    o1 = out_b[:1024]
    ms = timed(lambda: fe(wav[:1024], OUT_FRAMES, out=o1, per_clip_norm=True, return_n_frames=False), 20)
    out["per_clip_norm"] = dict(workload="1024 ESC-50 clips per GPU, kaldi recipe + per-clip mean / unbiased-std normalisation (2 kernels)",
                                ms_per_step=ms, value=world * 1024 * CLIP_SECONDS / (ms * 1e-3), unit=UNIT)
    # ---- Mixup fused into the epilogue vs frontend + stand-alone mixup kernel (every clip mixed, 2048-clip bank) ----
    NB = 2048
    bank = torch.randn((NB, OUT_FRAMES, N_MELS), generator=gen, device=dev)
    gm = torch.Generator().manual_seed(5 + rank)
    plan = b2.MixupPlan(torch.randint(0, NB, (1024,), generator=gm).int(), torch.rand(1024, generator=gm)).to(dev)

    def two_launches():
        fe(wav[:1024], OUT_FRAMES, mean=mean, std=std, out=o1, return_n_frames=False)
        b2.mixup_batch(o1, bank, plan, out=o1)
    ms2 = timed(two_launches, 20)
    ms1 = timed(lambda: fe(wav[:1024], OUT_FRAMES, mean=mean, std=std, out=o1, return_n_frames=False, mixup=(bank, plan)), 20)
    out["mixup_fused"] = dict(workload="1024 ESC-50 clips per GPU, every clip mixed with a partner from a 2048-clip bank in HBM: Mixup in the fbank "
                                       "epilogue (b200fbank_execute_mixup) vs frontend + stand-alone mixup kernel",
                              ms_per_step=ms1, ms_two_launches=ms2, value=world * 1024 * CLIP_SECONDS / (ms1 * 1e-3), unit=UNIT)
    del wav, out_b, bank
    # ---- the reference-actual recipe ------------------------------------
    Bm = 256
    wm = torch.rand((Bm, CLIP_SAMPLES), generator=gen, device=dev) * 2 - 1
    fem = b2.MelSpecFrontend(44100, 1024, 160, 400, N_MELS, 80.0, device=dev)
    ms = timed(lambda: fem(wm, out_frames=1379), 5)
    alg = Bm * (CLIP_SAMPLES * 4 + N_MELS * 1379 * 4)
    out["melspec"] = dict(workload="reference-actual recipe: 256 ESC-50 clips per GPU -> MelSpectrogram(1024/160 @44.1 kHz) + dB + per-clip "
                                   "normalisation -> (256,1,128,1379)",
                          ms_per_step=ms, value=world * Bm * CLIP_SECONDS / (ms * 1e-3), unit=UNIT,
                          roofline_frac=alg / (ms * 1e-3) / 1e9 / peak)
    del wm
    # ---- AST patch embedding fed by the frontend's features ---------------
    xs = torch.randn((1024, 1, N_MELS, OUT_FRAMES), generator=gen, device=dev)
    w16 = (torch.randn((768, 1, 16, 16), generator=gen, device=dev) * 0.05)
    pb = torch.zeros(768, device=dev)
    ms = timed(lambda: b2.patch_embed(xs, w16, pb, 10, torch.float16), 20)
    Mrows = 1024 * 12 * 50
    alg = 1024 * N_MELS * OUT_FRAMES * 4 + Mrows * 768 * 2 + 768 * 256 * 2
    out["patch_embed"] = dict(workload="1024 x (1,128,512) per GPU -> Conv2d(1,768,16,stride 10) + flatten -> (1024,600,768) fp16: TMA strips -> A in "
                                       "tensor memory -> tcgen05.mma (patch_embed_pipe_kernel)",
                              ms_per_step=ms, tflops=2.0 * Mrows * 768 * 256 / (ms * 1e-3) / 1e12,
                              roofline_frac=alg / (ms * 1e-3) / 1e9 / peak)
    return out


def run_extra(args, rank, local_rank, world):
    """Secondary BASELINE.json configs (documentation runs, not the driver's headline line):
    us8k  = configs[2]: 4096 ragged <=4 s clips at 22.05/44.1/48 kHz -> 1024 frames, SpecAugment masks;
    stats = configs[3]: per-bin sum / sum-of-squares over --clips synthetic clips sharded over the ranks + one
            all-reduce of 257 doubles;
    sweep = configs[4]: batch 64 .. 65536 clips (chunked at 4096 clips per launch)."""
    import random as _random
    import torch
    import torch.distributed as dist
    import dl_sound_classification_b200 as b2
    from dl_sound_classification_b200 import stats as ST

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup=3):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    out_lines = []
    if args.workload == "us8k":
        B, table = 4096, (22050, 44100, 48000)
        g = torch.Generator().manual_seed(31 + rank)
        rid = torch.randint(0, 3, (B,), generator=g)
        lens = ((1.0 + 3.0 * torch.rand(B, generator=g)) * torch.tensor(table)[rid]).long()
        offsets = torch.cat([torch.zeros(1, dtype=torch.int64), lens.cumsum(0)]).to(dev)
        flat = torch.rand(int(offsets[-1]), generator=gen, device=dev) * 2 - 1
        fe = b2.FbankFrontend(orig_rates=table, device=dev, **b2.AST_FBANK_KWARGS)
        _random.seed(77)
        masks = b2.specaugment.draw_masks(B, 1024, 128, 192, 48).to(dev)
        rid_d = rid.int().to(dev)
        out = torch.empty((B, 1024, N_MELS), device=dev)
        mean, std = torch.tensor([AST_MEAN], device=dev), torch.tensor([AST_STD], device=dev)
        ms = timed(lambda: fe(flat, 1024, offsets=offsets, rate_ids=rid_d, masks=masks, mean=mean, std=std, out=out,
                              return_n_frames=False), args.steps)
        secs = float((lens.double() / torch.tensor(table, dtype=torch.float64)[rid]).sum())
        alg = int(lens.sum()) * 4 + B * 1024 * N_MELS * 4
        out_lines.append(dict(workload="us8k: 4096 ragged clips (1-4 s @22.05/44.1/48 kHz) -> 1024 frames + SpecAugment masks",
                              ms_per_step=ms, value=world * secs / (ms * 1e-3), unit=UNIT,
                              roofline_frac=alg / (ms * 1e-3) / 1e9 / measured_peaks()[0], audio_seconds_per_step=secs))
    elif args.workload == "stats":
        total = args.clips
        lo, hi = ST.shard_bounds(total, rank, world)
        chunk = 4096
        fe = b2.FbankFrontend(orig_rates=(44100,), device=dev, **b2.AST_FBANK_KWARGS)
        ds = b2.DatasetStats(fe, OUT_FRAMES)
        wav = torch.empty((chunk, CLIP_SAMPLES), device=dev)
        ev = []
        barrier()
        for c0 in range(lo, hi, chunk):
            n = min(chunk, hi - c0)
            wav[:n].uniform_(-1, 1, generator=gen)                  # synthetic clips generated on the device (untimed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ds.update(wav[:n]); e1.record()
            ev.append((e0, e1))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ds.all_reduce(); e1.record()
        barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev), e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st = ds.finalize()
        ms = float(t[0]) + float(t[1])
        out_lines.append(dict(workload=f"stats: {total} synthetic ESC-50 clips sharded over {world} GPU(s), per-bin sum/sumsq + all-reduce",
                              ms_total=ms, ms_allreduce=float(t[1]), value=total * CLIP_SECONDS / (ms * 1e-3), unit=UNIT,
                              frames=st.frames, mean=st.mean, std=st.std,
                              roofline_frac=(total / world) * CLIP_SAMPLES * 4 / (float(t[0]) * 1e-3) / 1e9 / measured_peaks()[0]))
    elif args.workload == "sweep":
        fe = b2.FbankFrontend(orig_rates=(44100,), device=dev, **b2.AST_FBANK_KWARGS)
        mean, std = torch.tensor([AST_MEAN], device=dev), torch.tensor([AST_STD], device=dev)
        for B in (64, 256, 1024, 4096, 16384, 65536):
            nb = min(B, 4096)
            wav = torch.rand((nb, CLIP_SAMPLES), generator=gen, device=dev) * 2 - 1
            out = torch.empty((nb, OUT_FRAMES, N_MELS), device=dev)
            reps = B // nb

            def step():
                for _ in range(reps):                                  # > 4096 clips: streamed in 4096-clip launches
                    fe(wav, OUT_FRAMES, mean=mean, std=std, out=out, return_n_frames=False)
            ms = timed(step, max(3, args.steps // max(1, reps)))
            out_lines.append(dict(workload=f"sweep: batch {B} per GPU", ms_per_step=ms,
                                  value=world * B * CLIP_SECONDS / (ms * 1e-3), unit=UNIT,
                                  roofline_frac=B * (CLIP_SAMPLES * 4 + OUT_FRAMES * N_MELS * 4) / (ms * 1e-3) / 1e9 / measured_peaks()[0]))
            del wav, out
    elif args.workload == "mixup":
        # 1024 AST spectrograms (1, 128, 512) mixed with partners from a 2048-clip bank; every
        # sample mixed (2 reads + 1 write per element), so the algorithmic bytes are 3 x 268 MB
        B, NB = 1024, 2048
        x = torch.randn((B, 1, N_MELS, OUT_FRAMES), generator=gen, device=dev)
        bank = torch.randn((NB, 1, N_MELS, OUT_FRAMES), generator=gen, device=dev)
        g = torch.Generator().manual_seed(5 + rank)
        plan = b2.MixupPlan(torch.randint(0, NB, (B,), generator=g).int(), torch.rand(B, generator=g)).to(dev)
        out = torch.empty_like(x)
        ms = timed(lambda: b2.mixup_batch(x, bank, plan, out=out), args.steps)
        alg = 3 * B * N_MELS * OUT_FRAMES * 4
        out_lines.append(dict(workload="mixup: 1024 x (1,128,512) spectrograms, partners from a 2048-clip bank, all mixed",
                              ms_per_step=ms, value=world * B * CLIP_SECONDS / (ms * 1e-3), unit=UNIT,
                              roofline_frac=alg / (ms * 1e-3) / 1e9 / measured_peaks()[0], algorithmic_bytes_per_launch=alg))
    elif args.workload == "melspec":
        # the reference's own ASTPreprocessor recipe (MelSpectrogram n_fft 1024 / hop 160 at 44.1 kHz,
        # AmplitudeToDB top_db 80, per-clip mean / unbiased std normalisation) -> (B, 1, 128, 1379); generic kernel
        B = 256
        wav = torch.rand((B, CLIP_SAMPLES), generator=gen, device=dev) * 2 - 1
        fe = b2.MelSpecFrontend(44100, 1024, 160, 400, N_MELS, 80.0, device=dev)
        ms = timed(lambda: fe(wav, out_frames=1379), max(3, args.steps // 10))
        alg = B * (CLIP_SAMPLES * 4 + N_MELS * 1379 * 4)
        out_lines.append(dict(workload="melspec: reference-actual recipe, 256 ESC-50 clips -> (256,1,128,1379), dB + per-clip normalisation",
                              ms_per_step=ms, value=world * B * CLIP_SECONDS / (ms * 1e-3), unit=UNIT,
                              roofline_frac=alg / (ms * 1e-3) / 1e9 / measured_peaks()[0], algorithmic_bytes_per_launch=alg))
    elif args.workload == "patch_embed":
        # SURVEY.md section 8f N2 / BASELINE.json configs[4] tail: 1024 AST spectrograms (1, 128, 512) -> Conv2d(1, 768, 16,
        # stride 10) as a tcgen05 im2col GEMM -> (1024, 600, 768) fp16.  Algorithmic bytes: features in + embeddings out.
        B, D = 1024, 768
        x = torch.randn((B, 1, N_MELS, OUT_FRAMES), generator=gen, device=dev) * 0.5
        torch.manual_seed(7)
        conv = torch.nn.Conv2d(1, D, 16, stride=10).to(dev)
        w16, bias = conv.weight.detach(), conv.bias.detach()
        ms = timed(lambda: b2.patch_embed(x, w16, bias, 10, torch.float16), args.steps)
        npatch = 12 * ((OUT_FRAMES - 16) // 10 + 1)
        alg = B * N_MELS * OUT_FRAMES * 4 + B * npatch * D * 2 + D * 256 * 2
        flops = 2.0 * B * npatch * 256 * D
        out_lines.append(dict(workload="patch_embed: 1024 x (1,128,512) -> Conv2d(1,768,16,stride 10) -> (1024,600,768) fp16 (tcgen05)",
                              ms_per_step=ms, value=world * B * CLIP_SECONDS / (ms * 1e-3), unit=UNIT,
                              roofline_frac=alg / (ms * 1e-3) / 1e9 / measured_peaks()[0], algorithmic_bytes_per_launch=alg,
                              tflops=flops / (ms * 1e-3) / 1e12, dtype="f16"))
    if rank == 0:
        for d in out_lines:
            print(json.dumps(dict(dict(metric=METRIC, n_gpus=world, dtype="f32", data="synthetic"), **d)), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short configs[2..4] runs carried in the line's `extra` object")
    ap.add_argument("--workload", default="esc50", choices=["esc50", "us8k", "stats", "sweep", "mixup", "patch_embed", "melspec"],
                    help="esc50 = the headline line (BASELINE.json configs[1]); the others are documentation runs")
    ap.add_argument("--clips", type=int, default=100000, help="clips of the stats workload")
    ap.add_argument("--e2e-chunk", type=int, default=64, help="clips per pipelined chunk of the host-in/host-out path")
    args = ap.parse_args()
    if os.environ.get("B200_BENCH_WATCHDOG"):          # debugging aid: dump every thread's stack and exit if the run is still going after N s
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["B200_BENCH_WATCHDOG"]), exit=True)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload != "esc50":
        run_extra(args, rank, local_rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
