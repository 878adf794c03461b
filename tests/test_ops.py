"""torch.ops.b200fbank.*: schemas mirror the reference-side signatures; fake (meta) kernels give the right shapes; the
real kernels (GPU-marked) equal the package's Python entry points."""
import inspect

import pytest
import torch

import dl_sound_classification_b200 as b2


def test_kaldi_fbank_op_has_torchaudio_signature_and_defaults():
    K = pytest.importorskip("torchaudio.compliance.kaldi")
    ref = inspect.signature(K.fbank)
    ours = inspect.signature(b2.ops.kaldi_fbank._init_fn)
    assert [p.name for p in ref.parameters.values()] == [p.name for p in ours.parameters.values()]
    assert [p.default for p in ref.parameters.values()] == [p.default for p in ours.parameters.values()]
    schema = str(torch.ops.b200fbank.kaldi_fbank.default._schema)
    assert schema.startswith("b200fbank::kaldi_fbank(Tensor waveform, float blackman_coeff=") and "str window_type=\"povey\") -> Tensor" in schema


def test_fake_kernels_propagate_shapes():
    w = torch.empty(1, 80000, device="meta")
    assert tuple(torch.ops.b200fbank.kaldi_fbank(w, num_mel_bins=128, frame_shift=10.0).shape) == (498, 128)
    assert tuple(torch.ops.b200fbank.kaldi_fbank(w, num_mel_bins=40, use_energy=True, snip_edges=False).shape) == (500, 41)
    out, nfr = torch.ops.b200fbank.ast_frontend(torch.empty(7, 220500, device="meta"), 44100, 512, -6.6, 5.0)
    assert tuple(out.shape) == (7, 1, 128, 512) and tuple(nfr.shape) == (7,) and nfr.dtype == torch.int32
    x = torch.empty(3, 1, 128, 64, device="meta")
    assert tuple(torch.ops.b200fbank.mixup(x, torch.empty(9, 1, 128, 64, device="meta"), torch.empty(3, dtype=torch.int32, device="meta"),
                                           torch.empty(3, device="meta")).shape) == tuple(x.shape)


def test_ops_have_no_cpu_path():
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        torch.ops.b200fbank.mixup(torch.zeros(2, 4), torch.zeros(3, 4), torch.zeros(2, dtype=torch.int32), torch.ones(2))


@pytest.mark.gpu
def test_ops_equal_the_python_entry_points():
    from inputs import config1_clips, short_clip
    w = short_clip(24000)
    a = torch.ops.b200fbank.kaldi_fbank(w.cuda(), num_mel_bins=128, window_type="hanning", htk_compat=True, frame_shift=10.0)
    assert torch.equal(a, b2.fbank(w.cuda(), num_mel_bins=128, window_type="hanning", htk_compat=True, frame_shift=10.0))
    clips = torch.cat(config1_clips(3, length=50000), 0).cuda()
    out, nfr = torch.ops.b200fbank.ast_frontend(clips, 44100, 128, -6.6268, 5.0613)
    fe = b2.FbankFrontend(orig_rates=(44100,), **b2.AST_FBANK_KWARGS)
    ref, rn = fe(clips, out_frames=128, mean=-6.6268, std=5.0613, layout="bft")
    assert torch.equal(out, ref) and torch.equal(nfr, rn)
    plan = b2.MixupPlan(torch.tensor([2, -1, 0], dtype=torch.int32), torch.tensor([0.3, 1.0, 0.9]))
    assert torch.equal(torch.ops.b200fbank.mixup(out, ref, plan.partner.cuda(), plan.lam.cuda()), b2.mixup_batch(out, ref, plan))
