"""cavp_b200 - B200-native (sm_100a) implementation of the CAVP forward/backward hot path.

Public surface mirrors the reference (cyh-0/CAVP): `cavp_b200.models.cavp_model.CAVP` / `SoundBank`,
`cavp_b200.loss.ContrastLoss` / `CrossEntropyLoss` / `AVContrast`, `cavp_b200.optim.SGD` / `Adam` (fused optimiser steps),
`cavp_b200.metrics.MIoU` / `ForegroundDetect` (eval epilogue), `cavp_b200.audio.MelFrontEnd` (waveform -> log-mel), plus
`cavp_b200.trainer.train_step` (the restated trainer step running entirely on the C-ABI kernels).  There is no CPU or PyTorch fallback for the arithmetic."""
__version__ = "0.1.0"
