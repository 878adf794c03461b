"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol
include/b200fbank.h declares, its host-built tables and length arithmetic match the oracle
and the torchaudio golden fixtures, errors mirror torchaudio's, the SpecAugment replay is
bit-exact, and the stats all-reduce works over gloo with world_size 2.  No compute calls."""
import ctypes
import os
import random
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import dl_sound_classification_b200 as b2
from dl_sound_classification_b200 import _capi as K
from dl_sound_classification_b200 import specaugment as SA
from dl_sound_classification_b200 import stats as ST
from dl_sound_classification_b200.frontend import FbankFrontend, make_opts
from oracle import fbank_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200fbank.h")).read()
    declared = sorted(set(re.findall(r"\b(b200fbank_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    lib = ctypes.CDLL(K.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200fbank.h but not exported"
    assert sorted(K.EXPORTED_SYMBOLS) == declared
    assert lib.b200fbank_abi_version() == 1
    assert lib.b200fbank_sizeof_opts() == ctypes.sizeof(K.Opts)


def test_default_opts_match_torchaudio_defaults():
    o = K.default_opts()
    assert (o.blackman_coeff, o.energy_floor, o.frame_length, o.frame_shift) == (0.42, 1.0, 25.0, 10.0)
    assert (o.high_freq, o.low_freq, o.preemphasis_coefficient, o.sample_frequency) == (0.0, 20.0, 0.97, 16000.0)
    assert (o.vtln_high, o.vtln_low, o.vtln_warp, o.num_mel_bins) == (-500.0, 100.0, 1.0, 23)
    assert (o.window_type, o.htk_compat, o.raw_energy, o.remove_dc_offset) == (K.WINDOW_TYPES["povey"], 0, 1, 1)
    assert (o.round_to_power_of_two, o.snip_edges, o.subtract_mean, o.use_energy, o.use_log_fbank, o.use_power) == (1, 1, 0, 0, 1, 1)
    assert (o.lowpass_filter_width, o.rolloff) == (6, 0.99)


@pytest.fixture(scope="module")
def host_plan():
    return FbankFrontend(orig_rates=(44100, 22050, 48000, 16000), host_only=True, **b2.AST_FBANK_KWARGS)


def test_plan_tables_match_oracle_and_golden(host_plan, golden):
    p = host_plan.plan
    assert (p.window_shift, p.window_size, p.padded_window_size, p.n_cols) == (160, 400, 512, 128)
    for rid, (rate, width, L) in enumerate(((44100, 17, 34), (22050, 9, 17), (48000, 19, 37))):
        k, w, orig, new = O.sinc_resample_kernel(rate, 16000)
        info = p.rate_info(rid)
        assert (info["orig"], info["new"], info["width"], info["taps_per_phase"]) == (orig, new, width, L)
        t = p.table(K.TABLE_TAPS_DENSE, rid).reshape(k.shape)
        assert np.array_equal(t, k), f"taps for {rate} differ from the torchaudio recipe"
    g = golden("resample_kernels.npz")
    t = p.table(K.TABLE_TAPS_DENSE, 0).reshape(160, 475)
    assert np.array_equal(t[g["k44100_rows"]], g["k44100_rowvals"])           # bit-exact vs torchaudio
    assert np.array_equal(p.table(K.TABLE_WINDOW), O.kaldi_window("hanning", 400))
    mel = p.table(K.TABLE_MEL_DENSE).reshape(128, 256)
    gm = golden("mel_banks.npz")["b128"]
    assert np.abs(mel - gm).max() < 5e-5 and (mel != 0).sum() == 504 and (mel[3] == 0).all()
    assert np.abs(mel - O.kaldi_mel_banks(128, 512, 16000.0, 20.0, 0.0, dtype=np.float64)).max() < 1e-7


@pytest.mark.parametrize("window", ["povey", "hanning", "hamming", "rectangular", "blackman"])
def test_windows(window):
    fe = FbankFrontend(orig_rates=(16000,), host_only=True, window_type=window, blackman_coeff=0.4)
    got = fe.plan.table(K.TABLE_WINDOW)
    np.testing.assert_allclose(got, O.kaldi_window(window, 400, 0.4), rtol=0, atol=1e-7)


def test_vtln_and_other_banks(golden):
    fe = FbankFrontend(orig_rates=(16000,), host_only=True, num_mel_bins=40, high_freq=-400.0, vtln_warp=1.1)
    mel = fe.plan.table(K.TABLE_MEL_DENSE).reshape(40, 256)
    assert np.abs(mel - golden("mel_banks.npz")["b40_vtln"]).max() < 1e-4
    fe = FbankFrontend(orig_rates=(16000,), host_only=True)
    assert np.abs(fe.plan.table(K.TABLE_MEL_DENSE).reshape(23, 256) - golden("mel_banks.npz")["b23"]).max() < 5e-5


def test_length_arithmetic(host_plan):
    rng = random.Random(3)
    for rid, rate in enumerate((44100, 22050, 48000, 16000)):
        for n in [400, 441, 1103, 16000, 88200, 192000, 220500] + [rng.randrange(2000, 300000) for _ in range(50)]:
            n_rs = O.resampled_length(n, rate, 16000) if rate != 16000 else n
            assert host_plan.resampled_length(n, rid) == n_rs
            assert host_plan.num_frames(n, rid) == O.kaldi_num_frames(n_rs, 400, 160, True)
    assert host_plan.num_frames(220500, 0) == 498 and host_plan.num_frames(88200, 1) == 398
    fe = FbankFrontend(orig_rates=(16000,), host_only=True, snip_edges=False)
    assert fe.num_frames(24000) == 150
    with pytest.raises(ValueError):
        host_plan.rate_id(8000)


def test_option_errors_mirror_torchaudio():
    with pytest.raises(ValueError, match="preemphasis_coefficient"):
        FbankFrontend(host_only=True, preemphasis_coefficient=1.5)
    with pytest.raises(ValueError, match="at least 3 mel bins"):
        FbankFrontend(host_only=True, num_mel_bins=3)
    with pytest.raises(ValueError, match="Bad values in options"):
        FbankFrontend(host_only=True, low_freq=9000.0)
    with pytest.raises(ValueError, match="window_shift"):
        FbankFrontend(host_only=True, frame_shift=0.0)
    with pytest.raises(Exception, match="Invalid window type"):
        FbankFrontend(host_only=True, window_type="kaiser")
    with pytest.raises(NotImplementedError, match="power-of-two"):
        FbankFrontend(host_only=True, round_to_power_of_two=False)
    with pytest.raises(TypeError):
        make_opts((16000,), not_an_option=1)
    with pytest.raises(ValueError):
        FbankFrontend(orig_rates=tuple(range(8000, 8000 + 9)), host_only=True)


def test_no_cpu_compute_path(host_plan):
    w = torch.zeros(2, 2000)
    with pytest.raises(K.B200FbankError, match="no CPU compute path"):
        host_plan(w, out_frames=8)
    rc = K.lib.b200fbank_execute(host_plan.plan.handle, 1, None, 10, None, 1, None, None, None, 0, 0.0, 0.5, 8, 0, 1,
                                 None, None)
    assert rc == K.ERR_NO_DEVICE
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            b2.fbank(torch.zeros(1, 2000))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            b2.resample_waveform(torch.zeros(1, 2000), 44100, 16000)
        pre = b2.create_preprocessor("ast", dict(sample_rate=44100, n_mels=128), "/tmp/unused")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pre.preprocess(torch.zeros(1, 44100), 44100)
    assert b2.resample_waveform(w, 16000, 16000) is w                       # src/datasets/preprocessing.py:76
    with pytest.raises(NotImplementedError):
        b2.fbank(torch.zeros(1, 2000), dither=0.5)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dl_sound_classification_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "fbank_oracle" not in src or f.endswith(".inc"), f


def test_specaugment_replay_bit_exact(golden):
    g = golden("specaugment.npz")
    state = {}
    for F, T, s, draw, t0, tl, f0, fl in g["reference"]:
        key = (int(F), int(T), int(s))
        if draw == 0:
            state[key] = random.Random(int(s))
        assert SA.reference_intervals(int(T), int(F), 192, 48, state[key]) == (t0, tl, f0, fl)
    gens = {}
    for F, T, s, draw, t0, tl, f0, fl in g["torchaudio"]:
        key = (int(F), int(T), int(s))
        if draw == 0:
            gens[key] = torch.Generator().manual_seed(int(s))
        got = SA.torchaudio_intervals(int(T), int(F), 192, 48, gens[key])
        if tl == T or fl == F:
            continue
        assert got == (t0, tl, f0, fl), (key, draw, got)
    # product replay == independent oracle restatement, batch form
    random.seed(7)
    m = SA.draw_masks(5, [512, 512, 100, 1379, 300], 128)
    rng = O.PyRandom(7)
    want = [O.specaugment_intervals_reference(rng, t, 128) for t in (512, 512, 100, 1379, 300)]
    assert m.dtype == torch.int32 and m.tolist() == [list(w) for w in want]
    torch.manual_seed(11)
    m = SA.draw_masks(3, 512, 128, 192, 48, variant="torchaudio")
    gen = O.TorchCPUGenerator(11)
    assert m.tolist() == [list(O.specaugment_intervals_torchaudio(gen, 512, 128, 192, 48)) for _ in range(3)]
    # SURVEY.md 8c (6), (7)
    random.seed(42)
    pre = b2.ASTPreprocessor(b2.PreprocessingConfig(sample_rate=44100, n_mels=128))
    spec = torch.ones(1, 128, 1379)
    out = pre.apply_specaugment(spec, 192, 48)
    assert spec.eq(1).all() and out[0, :, 228:228 + 164].eq(0).all() and out[0, 94:96].eq(0).all()
    assert int(out.eq(0).sum()) == 164 * 128 + 2 * 1379 - 164 * 2
    torch.manual_seed(42)
    out = b2.SpecAugment(192, 48)(spec)
    assert out[0, :, 1106:1106 + 169].eq(0).all() and out[0, 105:105 + 18].eq(0).all()


def test_preprocessing_config_and_factory():
    cfg = b2.PreprocessingConfig(sample_rate=44100, n_mels=128, bc_mixing=False, normalize=True, target_mean=0.0,
                                 target_std=0.5)
    assert cfg.n_mels == 128 and cfg.config["target_std"] == 0.5 and len(cfg.get_hash()) == 12
    assert cfg.get_hash() != b2.PreprocessingConfig(sample_rate=44100, n_mels=64).get_hash()
    with pytest.raises(AttributeError):
        cfg.nope
    bad = b2.PreprocessingConfig(sample_rate=-1, n_mels="x")
    assert not bad.validate() and len(bad.get_validation_errors()) == 2
    pre = b2.create_preprocessor("ast", dict(sample_rate=44100, n_mels=128, target_frames=512, frontend="kaldi_fbank"),
                                 "/tmp/unused")
    assert isinstance(pre, b2.BasePreprocessor) and pre.get_cache_suffix().startswith("ast_")
    assert pre.config.config["target_frames"] == 512 and pre.target_sample_rate == 16000
    with pytest.raises(ValueError, match="Unknown preprocessing mode"):
        b2.create_preprocessor("nope", {}, "/tmp/unused")
    with pytest.raises(ValueError, match="Invalid preprocessing config"):
        b2.create_preprocessor("ast", dict(sample_rate=0), "/tmp/unused")
    with pytest.raises(NotImplementedError):
        b2.create_preprocessor("envnet_v2", {}, "/tmp/unused")


def test_stats_finalize_and_sharding(golden):
    sums = golden("config1.npz")["stats_sums"]
    st = ST.finalize_sums(torch.from_numpy(sums))
    mean_b, std_b, gm, gs = O.stats_finalize(sums)
    np.testing.assert_allclose(st.mean_per_bin.numpy(), mean_b, rtol=1e-12)
    np.testing.assert_allclose(st.std_per_bin.numpy(), std_b, rtol=1e-9, atol=1e-12)
    assert abs(st.mean - gm) < 1e-12 and abs(st.std - gs) < 1e-12 and st.frames == 40 * 498
    assert [ST.shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [ST.shard_bounds(100000, r, 8)[1] - ST.shard_bounds(100000, r, 8)[0] for r in range(8)] == [12500] * 8
    lens = [100, 100, 100, 900, 100, 100, 100, 100]
    b = [ST.shard_by_samples(lens, r, 2) for r in range(2)]
    assert b[0][0] == 0 and b[0][1] == b[1][0] and b[1][1] == 8 and 3 <= b[0][1] <= 4


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dl_sound_classification_b200 import stats as ST
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
lo, hi = ST.shard_bounds(10, rank, world)
sums = torch.zeros(257, dtype=torch.float64)
for i in range(lo, hi):                       # each "clip" contributes i to every sum, 498 frames
    sums[:256] += float(i)
    sums[256] += 498
ST.all_reduce_sums(sums)
assert sums[0].item() == sum(range(10)) and sums[256].item() == 4980, sums[[0, 256]]
st = ST.finalize_sums(sums)
assert st.frames == 4980
dist.destroy_process_group()
print("ok", rank)
"""


def test_stats_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_datamodule_mixin_draw_order_and_passthrough(monkeypatch):
    """CPU: the mixin consumes Python's `random` per sample in the reference's order (SpecAugment's four randints, then
    Mixup's coin / partner / coin), and leaves spectrogram batches alone.  The device work is stubbed out."""
    import random
    import types
    import dl_sound_classification_b200.datamodule as DMOD
    from dl_sound_classification_b200 import specaugment as SA

    seen = {}

    class StubPre:
        n_mels, sample_rate, target_frames = 128, 44100, 276
        frontend = None

        def preprocess_batch(self, wav, sr, masks=None, target_frames=None, mixup=None):
            seen["masks"], seen["plan"] = masks, mixup[1]
            return torch.zeros(wav.shape[0], 1, 128, target_frames), None

    monkeypatch.setattr(DMOD, "mixup_labels", lambda hard, bl, plan, n: torch.zeros(hard.shape[0], n))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))

    class DM(DMOD.B200DataModuleMixin):
        sample_rate, num_classes, time_mask, freq_mask, enable_mixup, mixup_alpha = 44100, 50, True, 48, True, 0.5
        trainer = types.SimpleNamespace(training=True)

    dm = DM()
    dm.setup_b200(preprocessor=StubPre())
    dm.set_mixup_bank(torch.zeros(9, 1, 128, 276), torch.arange(9))
    B = 5
    random.seed(11)
    torch.manual_seed(11)
    spec, soft = dm.on_after_batch_transfer((torch.zeros(B, 1, 1000), torch.arange(B)), 0)
    assert tuple(spec.shape) == (B, 1, 128, 276) and tuple(soft.shape) == (B, 50)
    random.seed(11)
    torch.manual_seed(11)
    rows, partners = [], []
    for _ in range(B):
        rows.append(list(SA.reference_intervals(276, 128, 192, 48, random)))
        p = -1
        if not random.random() > 0.5:
            other = random.randint(0, 8)
            if not random.random() > 0.5:
                p = other
                torch.distributions.Beta(0.5, 0.5).sample()
        partners.append(p)
    assert seen["masks"].tolist() == rows
    assert seen["plan"].partner.tolist() == partners
    b4 = (torch.zeros(2, 1, 128, 276), torch.zeros(2, 50))
    assert dm.on_after_batch_transfer(b4, 0) is b4
