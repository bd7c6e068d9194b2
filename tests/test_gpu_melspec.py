"""Mel front end on the GPU (csrc/pwv_mel.cuh through the C-ABI) against its CPU restatement (melspec.wav2melspec_db:
torch.stft on the CPU + the slaney basis pinned against torchaudio in test_melspec.py). librosa itself is not
installable here, so parity with the reference's librosa calls stays unpinned beyond those restated formulas."""
import numpy as np
import pytest

from conftest import pkg

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu


def _signals(n, t, sr, seed=0):
    rng = np.random.RandomState(seed)
    tt = np.arange(t) / sr
    out = []
    for i in range(n):
        f0 = 110.0 * (i + 1)
        x = sum(0.3 / (h + 1) * np.sin(2 * np.pi * f0 * (h + 1) * tt + rng.uniform(0, 6.28)) for h in range(8))
        x = x * (0.2 + 0.8 * np.abs(np.sin(2 * np.pi * 1.5 * tt))) + 0.01 * rng.randn(t)
        out.append(x)
    out[-1][: t // 3] = 0.0           # a silent stretch: exercises the amin floor and the top_db clamp
    return np.stack(out).astype(np.float32)


@pytest.mark.parametrize('sr,n_fft,win,hop,n_mels,t', [(16000, 512, 400, 80, 80, 16000), (24000, 512, 400, 80, 80, 4000),
                                                       (16000, 256, 256, 64, 40, 1000)])
def test_gpu_melspec_matches_cpu_restatement(hp, sr, n_fft, win, hop, n_mels, t):
    M = pkg('melspec')
    wav = _signals(3, t, sr)
    for norm in (True, False):
        kw = dict(max_db=hp.signal.max_db, min_db=hp.signal.min_db) if norm else {}
        ref = M.wav2melspec_db(wav, sr, n_fft, win, hop, n_mels, device='cpu', **kw).numpy()
        front = M.MelFrontEnd(sr, n_fft, win, hop, n_mels, **kw)
        got = front(torch.from_numpy(wav).cuda())
        torch.cuda.synchronize()
        got = got.cpu().numpy()
        assert got.shape == ref.shape == (3, 1 + t // hop, n_mels)
        # normalised range [-1, 1] spans 90 dB: 1e-4 there = 0.0045 dB; raw dB compared at 0.01 dB
        tol = 1e-4 if norm else 1e-2
        assert np.abs(got - ref).max() <= tol, (norm, float(np.abs(got - ref).max()))
        if norm:
            assert got.min() >= -1.0 and got.max() <= 1.0


def test_gpu_melspec_batch_independent_and_deterministic(hp):
    M = pkg('melspec')
    wav = torch.from_numpy(_signals(4, 8000, 16000, seed=3)).cuda()
    front = M.MelFrontEnd(16000, 512, 400, 80, 80, max_db=hp.signal.max_db, min_db=hp.signal.min_db)
    a = front(wav)
    b = front(wav)
    solo = front(wav[2:3])
    assert torch.equal(a, b) and torch.equal(solo[0], a[2])
    L = pkg('_lib')
    with pytest.raises(L.PwvError):
        front(wav[:, :200])             # shorter than the reflect padding
