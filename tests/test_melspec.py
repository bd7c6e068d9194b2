"""Mel front end (reference data_load.py:37-56 / audio.py) restated without librosa."""
import numpy as np
import pytest
import torch

from conftest import pkg


def test_mel_basis_matches_torchaudio_slaney():
    ta = pytest.importorskip('torchaudio')
    M = pkg('melspec')
    ours = M.mel_basis(16000, 512, 80)
    ref = ta.functional.melscale_fbanks(257, 0.0, 8000.0, 80, 16000, norm='slaney', mel_scale='slaney').T.numpy()
    assert ours.shape == (80, 257)
    assert np.abs(ours - ref).max() < 1e-6


def test_melspec_shape_range_and_known_tone(hp):
    M = pkg('melspec')
    sr, length = 16000, 4000
    t = np.arange(length) / sr
    wav = (0.5 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    mel = M.wav2melspec_db(wav, sr, 512, 400, 80, 80, max_db=hp.signal.max_db, min_db=hp.signal.min_db).numpy()
    assert mel.shape == (1 + length // 80, 80)                       # t_mel = 1 + T // hop (centred STFT)
    assert mel.min() >= -1.0 and mel.max() <= 1.0
    # the 1 kHz tone must peak in the mel band whose centre is nearest 1 kHz
    centres = M._mel_to_hz(np.linspace(M._hz_to_mel(0.0), M._hz_to_mel(8000.0), 82))[1:-1]
    assert abs(int(mel[25].argmax()) - int(np.abs(centres - 1000.0).argmin())) <= 1
    # amplitude_to_db semantics: unnormalised dB of a tone of amplitude 0.5 through a 400-sample hann window
    raw = M.wav2melspec_db(wav, sr, 512, 400, 80, 80).numpy()
    assert raw.max() - raw.min() <= 80.0 + 1e-3                      # top_db clipping
    silent = M.wav2melspec_db(np.zeros(length, np.float32), sr, 512, 400, 80, 80).numpy()
    assert np.allclose(silent, -100.0)                               # 20 log10(amin = 1e-5)


def test_trim_and_fix_length():
    M = pkg('melspec')
    sr = 16000
    wav = np.zeros(3 * sr, np.float32)
    wav[sr:2 * sr] = 0.3 * np.sin(2 * np.pi * 440 * np.arange(sr) / sr)
    trimmed = M.trim_wav(wav)
    assert sr <= len(trimmed) <= sr + 2 * 2048                       # silence removed up to one analysis frame per side
    assert len(M.fix_length(trimmed, 4000)) == 4000 and len(M.fix_length(trimmed[:100], 4000)) == 4000
    assert len(M.trim_wav(np.zeros(1000, np.float32))) == 1000    # librosa: ref = max -> all-zero input is kept


def test_generation_data_reads_wav_files(hp, tmp_path):
    from scipy.io import wavfile
    io = pkg('io')
    sr = 16000
    for i in range(12):
        t = np.arange(sr) / sr
        wavfile.write(str(tmp_path / f'utt{i:02d}.wav'), sr, (0.4 * np.sin(2 * np.pi * (200 + 50 * i) * t) * 32767).astype(np.int16))
    hp.set_hparam_dict({'data_path': str(tmp_path / '*.wav'), 'generate': {'batch_size': 2, 'length': 4000}}, case='wavtest')
    data = io.GenerationData(hp.data_path, hp.generate.batch_size, hp.generate.length)
    assert len(data.wav_files) == 12 - int(12 * 0.9)                 # reference data_load.py:22-23: the last 10 %
    gt, mel, noise = data.next_batch()
    assert gt.shape == (2, 4000, 1) and mel.shape == (2, 51, 80) and noise is None
    assert mel.dtype == np.float32 and -1.0 <= mel.min() and mel.max() <= 1.0


def test_write_wav_pcm16_round_trip(tmp_path):
    """io.write_wav (reference audio.py:19-20, PCM_16) -> melspec.read_wav: equal to within half an int16 step,
    clipped at full scale, sample rate kept."""
    import numpy as np
    from conftest import pkg
    io, M = pkg('io'), pkg('melspec')
    rng = np.random.RandomState(0)
    wav = np.concatenate([rng.uniform(-1, 1, 4000), [1.5, -1.5, 0.0, 32767.4 / 32768.0]]).astype(np.float32)
    path = str(tmp_path / 'x.wav')
    io.write_wav(wav, 16000, path)
    from scipy.io import wavfile
    sr, raw = wavfile.read(path)
    assert sr == 16000 and raw.dtype == np.int16 and raw.shape == (4004,)
    assert raw[4000] == 32767 and raw[4001] == -32768 and raw[4002] == 0
    back = M.read_wav(path, 16000)
    assert np.abs(back[:4000] - wav[:4000]).max() <= 0.5 / 32768 + 1e-7
