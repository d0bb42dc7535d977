"""Audio front-end: the restated oracle and the CUDA kernels against goldens produced by
torchaudio.transforms.MelSpectrogram + the unmodified utils/sourcesep.py (oracle/make_golden_mel.py).

Tolerance: the output is dB / 100 in [-1, 1]; 1e-3 relative (north_star) on values of order 1 = 0.1 dB.  Measured:
oracle vs torchaudio golden ~1e-6 (same torch.stft); CUDA (DFT as a 3xTF32 GEMM) ~1e-5."""
import pytest
import torch

from conftest import load_golden
from oracle import mel_oracle as MO

TOL = 1e-3


def _check(out, g):
    flat = out.detach().float().cpu().flatten()
    assert tuple(out.shape) == g["shape"]
    err = (flat[g["idx"]] - g["samples"]).abs().max().item()
    assert err < TOL, err
    assert abs(float(flat.double().sum()) - g["sum"]) < TOL * flat.numel() * 1e-2
    return err


def test_oracle_matches_torchaudio_golden():
    gold = load_golden("mel")
    fb = MO.melscale_fbanks_htk(257, 125.0, 3800.0, 64, 16000)
    # (tight allclose, not bit equality: the filterbank goes through pow / log10, whose last bit may differ between
    # vectorised and scalar code paths of different hosts)
    assert torch.allclose(fb, gold["fbanks"], rtol=1e-5, atol=1e-7)
    for g in gold["cases"]:
        c = g["case"]
        audio = MO.waveform_case(c["seed"], c["N"], c["C"], c["A"])
        out = MO.preprocess_audio(audio, audio_len=c["audio_len"])
        assert _check(out, g) < 1e-4


def test_host_constants_match_oracle():
    from cavp_b200.audio import dft_basis, hann_window_padded, melscale_fbanks_htk
    gold = load_golden("mel")
    assert torch.allclose(melscale_fbanks_htk(257, 125.0, 3800.0, 64, 16000), gold["fbanks"], rtol=1e-5, atol=1e-7)
    w = hann_window_padded(400, 512)
    assert w[:56].abs().sum() == 0 and w[456:].abs().sum() == 0 and torch.equal(w[56:456], torch.hann_window(400))
    # the basis reproduces torch.fft.rfft on a random frame
    x = torch.randn(512, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    b = dft_basis(512).double()
    ref = torch.fft.rfft(x)
    got = b @ x
    assert (got[:257] - ref.real).abs().max() < 2e-5 and (got[257:] - ref.imag).abs().max() < 2e-5


@pytest.mark.gpu
def test_cuda_frontend_matches_torchaudio_golden():
    from cavp_b200.audio import MelFrontEnd
    gold = load_golden("mel")
    for g in gold["cases"]:
        c = g["case"]
        audio = MO.waveform_case(c["seed"], c["N"], c["C"], c["A"])
        fe = MelFrontEnd(audio_len=c["audio_len"], spec_min=-100, spec_max=100)
        out = fe.preprocess_audio(audio.cuda())
        err = _check(out, g)
        ref = MO.preprocess_audio(audio, audio_len=c["audio_len"])
        full = (out.cpu() - ref).abs().max().item()
        print("mel front-end max abs err vs golden samples", err, "vs oracle (all elements)", full)
        assert full < TOL


@pytest.mark.gpu
def test_cuda_frontend_edge_cases():
    from cavp_b200 import _C
    from cavp_b200.audio import MelFrontEnd
    fe = MelFrontEnd(audio_len=1.0)
    silent = torch.zeros(1, 1, 16000).cuda()             # all-zero waveform: the 1e-5 floor -> 20*log10(1e-5) = -100 dB
    out = fe.preprocess_audio(silent)
    assert torch.allclose(out.cpu(), torch.full((1, 1, 96, 64), -1.0), atol=1e-6)
    ref = MO.preprocess_audio(torch.zeros(1, 1, 16000))
    assert torch.allclose(ref, torch.full((1, 1, 96, 64), -1.0), atol=1e-6)
    with pytest.raises(_C.CavpError):                    # waveform shorter than the reflect padding
        fe.preprocess_audio(torch.zeros(1, 1, 200).cuda())
    with pytest.raises(RuntimeError):
        fe.preprocess_audio(torch.zeros(1, 1, 16000))    # CPU tensor: no fallback
