"""Generate tests/golden/mel.pt with the reference's own front-end: torchaudio.transforms.MelSpectrogram configured as in
trainer/trainer_cavp_vpo_mono.py:43-53 and the UNMODIFIED utils/sourcesep.py (db_from_amp, normalize_spec), on CPU.

Build container only:   python oracle/make_golden_mel.py
"""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import mel_oracle as MO  # noqa: E402

CASES = [dict(seed=11, N=2, C=1, A=16000, audio_len=1.0), dict(seed=12, N=1, C=2, A=48000, audio_len=3.0)]
SAMPLES = 4096


def main():
    import torchaudio
    sys.path.insert(0, "/root/reference")
    from utils import sourcesep
    stft = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=512, win_length=400, hop_length=160, n_mels=64,
                                                f_min=125.0, f_max=3800.0)
    args = SimpleNamespace(spec_min=-100, spec_max=100)
    out = []
    for c in CASES:
        audio = MO.waveform_case(c["seed"], c["N"], c["C"], c["A"])
        N, C, A = audio.size()
        n_len = 96 if c["audio_len"] == 1.0 else 300
        a = audio.view(N * C, A)
        a = stft(a)[:, :, :n_len]
        a = a.transpose(-1, -2)
        a = sourcesep.db_from_amp(a, cuda=False)
        a = sourcesep.normalize_spec(a, args)
        _, T, F = a.size()
        a = a.view(N, C, T, F)
        flat = a.flatten()
        idx = (torch.arange(SAMPLES, dtype=torch.int64) * (flat.numel() - 1)) // (SAMPLES - 1)
        out.append(dict(case=c, shape=tuple(a.shape), idx=idx, samples=flat[idx].clone(), sum=float(flat.double().sum()),
                        min=float(flat.min()), max=float(flat.max())))
    path = os.path.join(ROOT, "tests", "golden", "mel.pt")
    torch.save(dict(torch_version=torch.__version__, torchaudio_version=torchaudio.__version__, cases=out,
                    fbanks=torchaudio.functional.melscale_fbanks(257, 125.0, 3800.0, 64, 16000)), path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
