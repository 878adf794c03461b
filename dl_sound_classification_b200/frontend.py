"""Batched B200 frontend: a plan (immutable tables on one GPU) + fused execute.

This is the host side of the C ABI in ``include/b200fbank.h``.  torch is used for
device memory, streams and the custom-op registration only; every stage of the
path runs in ``libb200fbank.so``.
"""
from __future__ import annotations

import threading
import weakref
from typing import Optional, Sequence, Tuple, Union

import torch

from . import _capi as K

_KALDI_DEFAULTS = dict(
    blackman_coeff=0.42, energy_floor=1.0, frame_length=25.0, frame_shift=10.0, high_freq=0.0,
    htk_compat=False, low_freq=20.0, num_mel_bins=23, preemphasis_coefficient=0.97, raw_energy=True,
    remove_dc_offset=True, round_to_power_of_two=True, sample_frequency=16000.0, snip_edges=True,
    subtract_mean=False, use_energy=False, use_log_fbank=True, use_power=True, vtln_high=-500.0,
    vtln_low=100.0, vtln_warp=1.0, window_type="povey",
)

AST_FBANK_KWARGS = dict(htk_compat=True, sample_frequency=16000.0, use_energy=False, window_type="hanning",
                        num_mel_bins=128, frame_shift=10.0)

# plan id -> frontend, for the torch custom operators (ops.py pass the id, not the object).  Weak references: a frontend
# (and its device tables) lives exactly as long as its owner keeps it.
_plans: "weakref.WeakValueDictionary[int, FbankFrontend]" = weakref.WeakValueDictionary()
_plans_lock = threading.Lock()
_next_plan_id = [1]


def _register(fe: "FbankFrontend") -> int:
    with _plans_lock:
        pid = _next_plan_id[0]
        _next_plan_id[0] += 1
        _plans[pid] = fe
    return pid


def _require_cuda(device) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("dl_sound_classification_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    d = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if d.type != "cuda":
        raise RuntimeError(f"device must be a CUDA device, got {d}")
    if d.index is None:
        d = torch.device("cuda", torch.cuda.current_device())
    return d


def make_opts(orig_rates: Sequence[int], lowpass_filter_width: int = 6, rolloff: float = 0.99, **kaldi) -> K.Opts:
    unknown = set(kaldi) - set(_KALDI_DEFAULTS)
    if unknown:
        raise TypeError(f"unknown kaldi.fbank option(s): {sorted(unknown)}")
    kw = dict(_KALDI_DEFAULTS)
    kw.update(kaldi)
    if kw["window_type"] not in K.WINDOW_TYPES:
        raise Exception("Invalid window type " + str(kw["window_type"]))      # torchaudio kaldi.py:113
    o = K.default_opts()
    for name in ("blackman_coeff", "energy_floor", "frame_length", "frame_shift", "high_freq", "low_freq",
                 "preemphasis_coefficient", "sample_frequency", "vtln_high", "vtln_low", "vtln_warp"):
        setattr(o, name, float(kw[name]))
    o.num_mel_bins = int(kw["num_mel_bins"])
    o.window_type = K.WINDOW_TYPES[kw["window_type"]]
    for name in ("htk_compat", "raw_energy", "remove_dc_offset", "round_to_power_of_two", "snip_edges",
                 "subtract_mean", "use_energy", "use_log_fbank", "use_power"):
        setattr(o, name, int(bool(kw[name])))
    rates = [int(r) for r in orig_rates]
    if not 1 <= len(rates) <= K.MAX_RATES:
        raise ValueError(f"between 1 and {K.MAX_RATES} input rates are supported")
    o.n_rates = len(rates)
    for i, r in enumerate(rates):
        o.orig_rates[i] = r
    o.lowpass_filter_width = int(lowpass_filter_width)
    o.rolloff = float(rolloff)
    return o


class FbankFrontend:
    """waveform batch -> (B, T, n_cols) or (B, 1, n_cols, T) log-mel features on one GPU.

    ``orig_rates`` is the table of input sample rates a clip may have
    (``rate_ids`` index it); the fbank itself runs at ``sample_frequency``.
    """

    def __init__(self, orig_rates: Sequence[int] = (16000,), device=None, lowpass_filter_width: int = 6,
                 rolloff: float = 0.99, host_only: bool = False, **kaldi):
        self.orig_rates = tuple(int(r) for r in orig_rates)
        self.kaldi = dict(_KALDI_DEFAULTS, **kaldi)
        opts = make_opts(self.orig_rates, lowpass_filter_width, rolloff, **kaldi)
        if host_only:
            self.device = None
            self.plan = K.Plan(opts, -1)
        else:
            self.device = _require_cuda(device)
            self.plan = K.Plan(opts, self.device.index)
        self.n_cols = self.plan.n_cols
        self.plan_id = _register(self)

    # -- host arithmetic ---------------------------------------------------------------
    def num_frames(self, n_samples: int, rate_id: int = 0) -> int:
        return self.plan.num_frames(int(n_samples), int(rate_id))

    def resampled_length(self, n_samples: int, rate_id: int = 0) -> int:
        return self.plan.resampled_length(int(n_samples), int(rate_id))

    def rate_id(self, sample_rate: int) -> int:
        try:
            return self.orig_rates.index(int(sample_rate))
        except ValueError:
            raise ValueError(f"sample rate {sample_rate} is not in this plan's rate table {self.orig_rates}") from None

    # -- argument plumbing -------------------------------------------------------------
    def _dev(self, t: Optional[torch.Tensor], dtype, what: str) -> Optional[torch.Tensor]:
        if t is None:
            return None
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        if t.device != self.device or t.dtype != dtype:
            t = t.to(device=self.device, dtype=dtype, non_blocking=True)
        return t.contiguous()

    def _wave_args(self, wav, offsets, rate_ids):
        if self.device is None:
            raise K.B200FbankError(K.ERR_NO_DEVICE, "host-only plan: no CPU compute path exists")
        if not wav.is_floating_point():
            raise TypeError(f"Expected floating point type for waveform tensor, but received {wav.dtype}.")
        wav = self._dev(wav, torch.float32, "wav")
        if offsets is None:
            if wav.dim() != 2:
                raise ValueError("dense batches must be (B, n_samples); pass offsets for ragged 1-D input")
            B, clip = int(wav.shape[0]), int(wav.shape[1])
            off = None
        else:
            if wav.dim() != 1:
                raise ValueError("ragged input must be a flat 1-D tensor with offsets[B+1]")
            off = self._dev(offsets, torch.int64, "offsets")
            B, clip = int(off.numel()) - 1, 0
        rid = self._dev(rate_ids, torch.int32, "rate_ids")
        if rid is not None and rid.numel() != B:
            raise ValueError("rate_ids must have one entry per clip")
        return wav, off, clip, rid, B

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]):
        return None if t is None else t.data_ptr()

    # -- device calls ------------------------------------------------------------------
    def __call__(self, wav: torch.Tensor, out_frames: int, offsets: Optional[torch.Tensor] = None,
                 rate_ids: Optional[torch.Tensor] = None, masks: Optional[torch.Tensor] = None,
                 mean: Union[None, float, torch.Tensor] = None, std: Union[None, float, torch.Tensor] = None,
                 target_mean: float = 0.0, target_std: float = 0.5, layout: str = "btf",
                 out: Optional[torch.Tensor] = None, return_n_frames: bool = True, per_clip_norm: bool = False,
                 remove_clip_mean: bool = False, mixup=None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """The fused path (``b200fbank_execute``).  ``mean``/``std``: None (no
        normalisation), scalar, or per-column ``[n_cols]``; output
        ``(x-mean)/std*target_std+target_mean``.  ``masks``: int32 ``(B, 4)`` =
        ``t_start, t_len, f_start, f_len``; cells zeroed after normalisation.
        ``per_clip_norm``: every clip is normalised with its OWN mean / unbiased std over its real frames instead
        (``b200fbank_clip_normalize``, the reference's src/datasets/preprocessing.py:1030-1037), masks after it.
        ``remove_clip_mean``: ``waveform - waveform.mean()`` per clip before the resampler
        (``b200fbank_remove_clip_mean``; the AST recipe's convention).
        ``mixup``: ``(bank, plan)`` -- Mixup fused into the kernel's epilogue (``b200fbank_execute_mixup``): ``bank`` holds
        spectrograms shaped like the output (same layout and ``out_frames``), ``plan`` is a ``mixup.MixupPlan``; the result
        is bit-identical to ``mixup.mixup_batch`` applied to the un-mixed output, without the extra pass over the batch."""
        wav, off, clip, rid, B = self._wave_args(wav, offsets, rate_ids)
        if per_clip_norm and mean is not None:
            raise ValueError("per_clip_norm uses each clip's own statistics: do not pass mean/std")
        if remove_clip_mean and B > 0:
            centred = torch.empty_like(wav)
            with torch.cuda.device(self.device):
                K.check(K.lib.b200fbank_remove_clip_mean(wav.data_ptr(), self._ptr(off), clip, B, centred.data_ptr(), None,
                                                         torch.cuda.current_stream(self.device).cuda_stream))
            wav = centred
        lay = {"btf": K.LAYOUT_BTF, "bft": K.LAYOUT_BFT}[layout]
        shape = (B, int(out_frames), self.n_cols) if lay == K.LAYOUT_BTF else (B, 1, self.n_cols, int(out_frames))
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or out.device != self.device or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 {shape} tensor on {self.device}")
        nfr = torch.empty(B, dtype=torch.int32, device=self.device) if (return_n_frames or per_clip_norm) else None
        mk = self._dev(masks, torch.int32, "masks")
        if mk is not None and tuple(mk.shape) != (B, 4):
            raise ValueError("masks must be (B, 4) int32: t_start, t_len, f_start, f_len")
        if (mean is None) != (std is None):
            raise ValueError("pass both mean and std or neither")
        mt = st = None
        n_stats = 0
        if mean is not None:
            mt = self._dev(torch.as_tensor(mean, dtype=torch.float32).reshape(-1), torch.float32, "mean")
            st = self._dev(torch.as_tensor(std, dtype=torch.float32).reshape(-1), torch.float32, "std")
            if mt.numel() != st.numel():
                raise ValueError("mean and std must have the same number of elements")
            n_stats = int(mt.numel())
        bank = partner = lam = None
        if mixup is not None:
            bank, plan = mixup
            if bank.dtype != torch.float32 or tuple(bank.shape[1:]) != shape[1:]:
                raise ValueError(f"the mixup bank must be float32 (N, {', '.join(map(str, shape[1:]))}), got {tuple(bank.shape)}")
            bank = self._dev(bank, torch.float32, "mixup bank")
            partner = self._dev(plan.partner, torch.int32, "mixup partner")
            lam = self._dev(plan.lam, torch.float32, "mixup lam")
            if int(partner.numel()) != B or int(lam.numel()) != B:
                raise ValueError("the mixup plan does not match the batch")
            if B and int(partner.max()) >= bank.shape[0]:
                raise IndexError("partner index outside the bank")
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            fused_mix = bank is not None and not per_clip_norm
            head = (self.plan.handle, wav.data_ptr(), self._ptr(off), clip, self._ptr(rid), B,
                    None if per_clip_norm else self._ptr(mk),
                    self._ptr(mt), self._ptr(st), n_stats, float(target_mean), float(target_std), int(out_frames), lay)
            if fused_mix:
                K.check(K.lib.b200fbank_execute_mixup(*head, bank.data_ptr(), partner.data_ptr(), lam.data_ptr(),
                                                      out.data_ptr(), self._ptr(nfr), stream))
            else:
                K.check(K.lib.b200fbank_execute(*head, out.data_ptr(), self._ptr(nfr), stream))
            if per_clip_norm and B > 0:
                K.check(K.lib.b200fbank_clip_normalize(out.data_ptr(), nfr.data_ptr(), B, int(out_frames), self.n_cols, lay,
                                                       None, -1.0, 1, float(target_mean), float(target_std), self._ptr(mk), stream))
                if bank is not None:       # per-clip statistics need the whole clip first: Mixup follows as its own launch
                    K.check(K.lib.b200fbank_mixup(out.data_ptr(), bank.data_ptr(), partner.data_ptr(), lam.data_ptr(), B,
                                                  int(out_frames) * self.n_cols, out.data_ptr(), stream))
        return out, (nfr if return_n_frames else None)

    def pcm16_to_float(self, pcm: torch.Tensor, divisor: Optional[torch.Tensor] = None,
                       offsets: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``b200fbank_pcm16_to_float``: int16 PCM on the device -> float32 ``pcm / divisor[clip]`` (default 32768 =
        ``torchaudio.load``; the clip's peak = the normalisation of scripts/prepare_esc50.py:94-101, bit for bit)."""
        if pcm.dtype != torch.int16:
            raise TypeError("pcm must be int16")
        pcm = pcm.to(self.device).contiguous()
        off = self._dev(offsets, torch.int64, "offsets")
        if off is None:
            if pcm.dim() != 2:
                raise ValueError("dense batches must be (B, n_samples); pass offsets for ragged 1-D input")
            B, clip, longest = int(pcm.shape[0]), int(pcm.shape[1]), 0
        else:
            B, clip = int(off.numel()) - 1, 0
            longest = int((off[1:] - off[:-1]).max()) if B > 0 else 0
        dv = self._dev(divisor, torch.float32, "divisor")
        if dv is not None and dv.numel() != B:
            raise ValueError("divisor must have one entry per clip")
        if out is None:
            out = torch.empty(pcm.shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            K.check(K.lib.b200fbank_pcm16_to_float(pcm.data_ptr(), self._ptr(off), clip, B, self._ptr(dv), longest,
                                                   out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def process_host(self, h_wav: torch.Tensor, out_frames: int, h_out: Optional[torch.Tensor] = None,
                     chunk_clips: int = 64, n_streams: int = 3, layout: str = "btf",
                     pcm_divisor: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
        """Host buffers in, host buffers out: the end-to-end form of the fused path.

        ``h_wav`` is a dense ``(B, n_samples)`` CPU tensor (pinned memory for full PCIe speed): float32 waveforms, or
        int16 PCM (half the host -> device bytes; widened on the device as ``pcm / pcm_divisor[clip]``, default 32768,
        see ``pcm16_to_float``).  The result is a CPU tensor (pinned when allocated here).  The batch is cut into
        chunks that rotate over ``n_streams`` CUDA streams, so the host->device copy of chunk
        i+1, the kernel of chunk i and the device->host copy of chunk i-1 overlap (PCIe is full
        duplex; the kernel is ~30x faster than either copy).  Keyword arguments are those of
        ``__call__`` (``mean``, ``std``, ``target_mean``, ``target_std``); per-clip ``masks`` are
        sliced per chunk.
        """
        if self.device is None:
            raise K.B200FbankError(K.ERR_NO_DEVICE, "host-only plan: no CPU compute path exists")
        if h_wav.device.type != "cpu" or h_wav.dim() != 2 or h_wav.dtype not in (torch.float32, torch.int16):
            raise ValueError("h_wav must be a dense (B, n_samples) float32 or int16 CPU tensor")
        pcm = h_wav.dtype == torch.int16
        div = self._dev(pcm_divisor, torch.float32, "pcm_divisor") if pcm else None
        B, N = int(h_wav.shape[0]), int(h_wav.shape[1])
        shape = (B, int(out_frames), self.n_cols) if layout == "btf" else (B, 1, self.n_cols, int(out_frames))
        if h_out is None:
            h_out = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        elif tuple(h_out.shape) != shape or h_out.dtype != torch.float32 or h_out.device.type != "cpu":
            raise ValueError(f"h_out must be a float32 CPU tensor of shape {shape}")
        masks = kw.pop("masks", None)
        for k in ("mean", "std"):                       # upload the statistics once, not per chunk
            if kw.get(k) is not None:
                kw[k] = self._dev(torch.as_tensor(kw[k], dtype=torch.float32).reshape(-1), torch.float32, k)
        chunk_clips = max(1, min(int(chunk_clips), B))
        state = getattr(self, "_host_pipe", None)
        key = (chunk_clips, N, shape[1:], n_streams, pcm)
        if state is None or state["key"] != key:
            state = dict(key=key, streams=[torch.cuda.Stream(self.device) for _ in range(n_streams)],
                         d_wav=[torch.empty((chunk_clips, N), dtype=torch.float32, device=self.device) for _ in range(n_streams)],
                         d_pcm=[torch.empty((chunk_clips, N), dtype=torch.int16, device=self.device) for _ in range(n_streams)]
                         if pcm else None,
                         d_out=[torch.empty((chunk_clips,) + shape[1:], dtype=torch.float32, device=self.device)
                                for _ in range(n_streams)])
            self._host_pipe = state
        cur = torch.cuda.current_stream(self.device)
        for st in state["streams"]:
            st.wait_stream(cur)
        for i, lo in enumerate(range(0, B, chunk_clips)):
            hi = min(B, lo + chunk_clips)
            n = hi - lo
            k = i % n_streams
            with torch.cuda.stream(state["streams"][k]):
                dw, do = state["d_wav"][k][:n], state["d_out"][k][:n]
                if pcm:
                    dp = state["d_pcm"][k][:n]
                    dp.copy_(h_wav[lo:hi], non_blocking=True)
                    self.pcm16_to_float(dp, None if div is None else div[lo:hi], out=dw)
                else:
                    dw.copy_(h_wav[lo:hi], non_blocking=True)
                self(dw, out_frames, masks=None if masks is None else masks[lo:hi], layout=layout, out=do,
                     return_n_frames=False, **kw)
                h_out[lo:hi].copy_(do, non_blocking=True)
        for st in state["streams"]:
            cur.wait_stream(st)
            st.synchronize()        # "host buffers out": the D2H copies have LANDED when the caller gets h_out back
        return h_out

    def resample(self, wav: torch.Tensor, offsets: Optional[torch.Tensor] = None,
                 rate_ids: Optional[torch.Tensor] = None, out_clip_samples: Optional[int] = None) -> torch.Tensor:
        """``b200fbank_resample``: dense ``(B, n)`` -> ``(B, ceil(new*n/orig))``.  With
        ``rate_ids``/ragged input the rows are ``out_clip_samples`` long (zero beyond
        each clip's own resampled length)."""
        wav, off, clip, rid, B = self._wave_args(wav, offsets, rate_ids)
        if out_clip_samples is None:
            if off is not None or rid is not None:
                raise ValueError("out_clip_samples is required for ragged / mixed-rate input")
            out_clip_samples = self.resampled_length(clip, 0)
        out = torch.zeros((B, int(out_clip_samples)), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            K.check(K.lib.b200fbank_resample(self.plan.handle, wav.data_ptr(), self._ptr(off), clip, self._ptr(rid), B,
                                             out.data_ptr(), None, int(out_clip_samples), stream))
        return out

    def accumulate_stats(self, wav: torch.Tensor, sums: torch.Tensor, max_frames: int,
                         offsets: Optional[torch.Tensor] = None, rate_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``b200fbank_stats_accumulate``: add this batch's per-column sum / sum of
        squares / frame count to ``sums`` (float64 ``[2*n_cols+1]`` on the GPU)."""
        wav, off, clip, rid, B = self._wave_args(wav, offsets, rate_ids)
        if sums.dtype != torch.float64 or sums.device != self.device or sums.numel() != 2 * self.n_cols + 1 \
                or not sums.is_contiguous():
            raise ValueError(f"sums must be a contiguous float64 [{2 * self.n_cols + 1}] tensor on {self.device}")
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            K.check(K.lib.b200fbank_stats_accumulate(self.plan.handle, wav.data_ptr(), self._ptr(off), clip,
                                                     self._ptr(rid), B, int(max_frames), sums.data_ptr(), stream))
        return sums


class MelSpecFrontend(FbankFrontend):
    """The reference-ACTUAL recipe: ``MelSpectrogram(sample_rate, n_fft, hop_length,
    win_length, n_mels, power=2)`` + ``AmplitudeToDB(top_db)`` + the per-clip normalisation of
    ``ASTPreprocessor.preprocess`` (src/datasets/preprocessing.py:988-998, 1013-1039; src/utils/audio.py:60-84),
    batched on the GPU.  Features are computed at ``sample_rate``; clips at another rate of ``orig_rates`` are
    resampled first (``resample_waveform``, :61-76)."""

    def __init__(self, sample_rate: int = 44100, n_fft: int = 1024, hop_length: int = 160,
                 win_length: Optional[int] = 400, n_mels: int = 128, top_db: Optional[float] = 80.0,
                 orig_rates: Optional[Sequence[int]] = None, device=None, host_only: bool = False):
        self.orig_rates = tuple(int(r) for r in (orig_rates or (sample_rate,)))
        self.kaldi = {}
        opts = make_opts(self.orig_rates, sample_frequency=float(sample_rate), num_mel_bins=int(n_mels),
                         low_freq=0.0, high_freq=0.0)
        opts.frontend = K.FRONTEND_MELSPEC_DB
        opts.n_fft, opts.hop_length = int(n_fft), int(hop_length)
        opts.win_length = int(win_length) if win_length else int(n_fft)
        opts.top_db = float(top_db) if top_db is not None else -1.0
        self.sample_rate, self.top_db = int(sample_rate), top_db
        if host_only:
            self.device = None
            self.plan = K.Plan(opts, -1)
        else:
            self.device = _require_cuda(device)
            self.plan = K.Plan(opts, self.device.index)
        self.n_cols = self.plan.n_cols
        self.plan_id = _register(self)

    def __call__(self, wav: torch.Tensor, out_frames: int, offsets: Optional[torch.Tensor] = None,
                 rate_ids: Optional[torch.Tensor] = None, masks: Optional[torch.Tensor] = None, to_db: bool = True,
                 normalize: bool = True, target_mean: float = 0.0, target_std: float = 0.5, layout: str = "bft",
                 out: Optional[torch.Tensor] = None, return_n_frames: bool = True):
        wav, off, clip, rid, B = self._wave_args(wav, offsets, rate_ids)
        lay = {"btf": K.LAYOUT_BTF, "bft": K.LAYOUT_BFT}[layout]
        shape = (B, int(out_frames), self.n_cols) if lay == K.LAYOUT_BTF else (B, 1, self.n_cols, int(out_frames))
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or out.device != self.device or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous float32 {shape} tensor on {self.device}")
        nfr = torch.empty(max(B, 1), dtype=torch.int32, device=self.device)[:B]      # the per-clip pass reads it
        mk = self._dev(masks, torch.int32, "masks")
        if mk is not None and tuple(mk.shape) != (B, 4):
            raise ValueError("masks must be (B, 4) int32: t_start, t_len, f_start, f_len")
        cmax = torch.empty(max(B, 1), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            K.check(K.lib.b200fbank_melspec_db(
                self.plan.handle, wav.data_ptr(), self._ptr(off), clip, self._ptr(rid), B, self._ptr(mk),
                int(bool(to_db)), int(bool(normalize)), float(target_mean), float(target_std), int(out_frames), lay,
                out.data_ptr(), nfr.data_ptr(), cmax.data_ptr(), stream))
        return out, (nfr if return_n_frames else None)


def launch_count(reset: bool = False) -> int:
    """Kernel launches issued by this thread through the C ABI (bench.py ``gpu_launches``)."""
    return int(K.lib.b200fbank_launch_count(1 if reset else 0))


def get_plan(plan_id: int) -> FbankFrontend:
    return _plans[plan_id]
