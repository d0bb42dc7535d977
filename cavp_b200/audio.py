"""Audio front-end on device: waveform -> normalised log-mel, the input of the audio backbone.

Drop-in for `CAVP_TRAINER.preprocess_audio` (trainer/trainer_cavp_vpo_mono.py:43-53,61-71), i.e.

    torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=64,
                                         f_min=125.0, f_max=3800.0)(audio)[:, :, :n_len].transpose(-1, -2)
    -> sourcesep.db_from_amp -> sourcesep.normalize_spec(spec_min, spec_max)            (utils/sourcesep.py:23-47)

on three kernels: framing + window (csrc/audio.cu), the real DFT as one tcgen05 GEMM against a constant cos | -sin
basis (cavp_igemm, fp32-parity mode), power -> mel filterbank -> dB -> range (csrc/audio.cu).  The window, DFT basis
and the HTK triangular filterbank (torchaudio.functional.melscale_fbanks, restated) are constants built once on the
host.  No CPU fallback.
"""
import math

import torch

from . import _C


def hann_window_padded(win_length, n_fft):
    """torch.hann_window(win_length) (periodic) zero-padded to n_fft on both sides, as torch.stft does."""
    w = torch.hann_window(win_length, periodic=True, dtype=torch.float32)
    left = (n_fft - win_length) // 2
    out = torch.zeros(n_fft, dtype=torch.float32)
    out[left:left + win_length] = w
    return out


def melscale_fbanks_htk(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") restated with the same fp32 tensor ops
    (torchaudio/functional/functional.py: _hz_to_mel, _mel_to_hz, _create_triangular_filterbank)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down_slopes, up_slopes))  # [n_freqs, n_mels]


def dft_basis(n_fft):
    """[2*(n_fft/2+1), n_fft] rows: cos(2 pi k j / n) for k = 0..n/2, then -sin(...) (X_k = sum_j x_j e^{-i 2 pi k j/n})."""
    nf = n_fft // 2 + 1
    k = torch.arange(nf, dtype=torch.float64).unsqueeze(1)
    j = torch.arange(n_fft, dtype=torch.float64).unsqueeze(0)
    ang = 2.0 * math.pi * ((k * j) % n_fft) / n_fft
    return torch.cat((torch.cos(ang), -torch.sin(ang)), 0).float().contiguous()


class MelFrontEnd:
    """preprocess_audio(audio [N, C, A]) -> [N, C, T, n_mels], T = 96 if audio_len == 1.0 else 300."""

    def __init__(self, args=None, sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=64, f_min=125.0,
                 f_max=3800.0, spec_min=None, spec_max=None, audio_len=None, prec=2):
        self.n_fft, self.hop, self.n_mels = n_fft, hop_length, n_mels
        self.nf = n_fft // 2 + 1
        self.spec_min = float(spec_min if spec_min is not None else getattr(args, "spec_min", -100))
        self.spec_max = float(spec_max if spec_max is not None else getattr(args, "spec_max", 100))
        self.audio_len = float(audio_len if audio_len is not None else getattr(args, "audio_len", 1.0))
        self.prec = prec
        self._window = hann_window_padded(win_length, n_fft)
        self._fb = melscale_fbanks_htk(self.nf, f_min, f_max, n_mels, sample_rate).contiguous()
        self._basis = dft_basis(n_fft)
        self._dev = {}

    def _consts(self, dev):
        c = self._dev.get(dev)
        if c is None:
            basis = self._basis.to(dev)
            split = torch.empty(2, *basis.shape, device=dev)
            _C.call("cavp_split_tf32", basis.data_ptr(), split[0].data_ptr(), split[1].data_ptr(), basis.numel(),
                    torch.cuda.current_stream(dev).cuda_stream)
            c = dict(window=self._window.to(dev), fb=self._fb.to(dev), split=split)
            self._dev[dev] = c
        return c

    def n_frames(self):
        return 96 if self.audio_len == 1.0 else 300

    def preprocess_audio(self, audio):
        if not audio.is_cuda:
            raise RuntimeError("cavp_b200.audio runs on CUDA (sm_100a) only; there is no CPU fallback")
        N, C, A = audio.size()
        dev = audio.device
        c = self._consts(dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        wave = audio.reshape(N * C, A).float().contiguous()
        rows, T, n_fft, nf = N * C, self.n_frames(), self.n_fft, self.nf
        frames = torch.empty(rows * T, n_fft, device=dev)
        _C.call("cavp_mel_frames", wave.data_ptr(), A, A, rows, T, n_fft, self.hop, c["window"].data_ptr(),
                frames.data_ptr(), st)
        ncols = 2 * nf
        lds = (ncols + 3) // 4 * 4
        spec = torch.empty(rows * T, lds, device=dev)
        # real DFT: [rows*T, n_fft] x basis^T, a Linear in cavp_igemm terms (nimg = rows*T, 1x1 "image", r = s = 1)
        _C.call("cavp_igemm", frames.data_ptr(), c["split"][0].data_ptr(), spec.data_ptr(), 0, 0, 0, 0, 0, rows * T, 1, 1,
                n_fft, n_fft, 1, 1, 1, 1, 1, 0, 1, 0, ncols, n_fft, lds, 0, 0, 0, 0, 0, 0.0, 1, self.prec,
                c["split"][0].numel(), st)
        out = torch.empty(N, C, T, self.n_mels, device=dev)
        _C.call("cavp_mel_power_db", spec.data_ptr(), lds, nf, c["fb"].data_ptr(), self.n_mels, rows * T, 1e-5, 20.0,
                self.spec_min, self.spec_max, out.data_ptr(), st)
        return out

    __call__ = preprocess_audio
