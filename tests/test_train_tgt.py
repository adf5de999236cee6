"""Training-target row: mixing, instantaneous a priori SNR, mapped target and the mergeable per-bin
statistics of xi_dB.  CPU tests: the oracle against the reference's formulation (mean / std of the stacked sample) and
the world_size-2 gloo all-reduce of the moments; GPU tests: the CUDA path (dxi_mix, dxi_xi_map, dxi_xi_db_moments behind
MagXi.mix / example / stats) against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepxi_b200 import stats
from oracle import cdfmap, sig, train_tgt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _corpus(n=6, seed=3):
    rng = np.random.default_rng(seed)
    s_len = [int(v) for v in rng.integers(3000, 9000, n)]
    d_len = [int(l + rng.integers(0, 5000)) for l in s_len]
    Ls, Ld = max(s_len), max(d_len)
    s = np.zeros((n, Ls), np.int16)
    d = np.zeros((n, Ld), np.int16)
    for i in range(n):
        a = np.convolve(rng.standard_normal(s_len[i] + 63), 0.9 ** np.arange(64), mode='valid')
        s[i, :s_len[i]] = np.clip(a / np.abs(a).max() * 12000, -32768, 32767).astype(np.int16)
        d[i, :d_len[i]] = np.clip(rng.standard_normal(d_len[i]) * 2500, -32768, 32767).astype(np.int16)
    snr = [float(v) for v in rng.integers(-10, 21, n)]
    off = [int(rng.integers(0, 1 + dl - sl)) for sl, dl in zip(s_len, d_len)]
    return s, d, s_len, d_len, snr, off


def test_oracle_mix_hits_the_snr_and_matches_the_stacked_statistics():
    s, d, s_len, d_len, snr, off = _corpus()
    so, do, xo, nfr = train_tgt.mix(s, d, s_len, d_len, snr, off)
    for i, n in enumerate(s_len):
        got = 10 * np.log10(np.mean(so[i, :n].astype(np.float64) ** 2) / np.mean(do[i, :n].astype(np.float64) ** 2))
        assert abs(got - snr[i]) < 1e-3
        assert np.array_equal(xo[i, :n], so[i, :n] + do[i, :n]) and not xo[i, n:].any()
        assert nfr[i] == -(-n // 256)
    # the reference stacks the frames of all utterances and takes mean / std per bin (inp_tgt.py:126-139, map.py:401-402)
    rows = []
    for i, n in enumerate(s_len):
        S, _ = sig.polar_analysis(so[i, :n])
        D, _ = sig.polar_analysis(do[i, :n])
        rows.append(cdfmap.db(train_tgt.xi(S, D)))
    stacked = np.vstack(rows).astype(np.float64)
    mu, sg = train_tgt.stats_from_moments(train_tgt.xi_db_moments(so, do, s_len))
    assert np.abs(mu - stacked.mean(axis=0)).max() < 1e-4 and np.abs(sg - stacked.std(axis=0)).max() < 1e-4
    mu2, sg2 = stats.stats_from_moments(train_tgt.xi_db_moments(so, do, s_len))
    assert np.array_equal(mu, mu2) and np.array_equal(sg, sg2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    s, d, s_len, d_len, snr, off = _corpus()
    so, do, _, _ = train_tgt.mix(s, d, s_len, d_len, snr, off)
    mine = list(range(rank, len(s_len), world))                      # this rank's shard of the sample
    acc = torch.from_numpy(train_tgt.xi_db_moments(so[mine], do[mine], [s_len[i] for i in mine]))
    acc = stats.allreduce_moments(acc)                               # the collective of the stats path
    if rank == 0:
        q.put(acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_of_the_moments_equals_the_whole_sample():
    s, d, s_len, d_len, snr, off = _corpus()
    so, do, _, _ = train_tgt.mix(s, d, s_len, d_len, snr, off)
    whole = train_tgt.xi_db_moments(so, do, s_len)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(merged[0], whole[0])
    assert np.allclose(merged, whole, rtol=1e-12, atol=0)
    mu_a, sg_a = stats.stats_from_moments(merged)
    mu_b, sg_b = stats.stats_from_moments(whole)
    assert np.abs(mu_a - mu_b).max() < 1e-5 and np.abs(sg_a - sg_b).max() < 1e-5


# ---- GPU -------------------------------------------------------------------------------------------------------
def _it():
    from deepxi_b200.inp_tgt import inp_tgt_selector
    it = inp_tgt_selector('MagXi', 512, 256, 512, 16000, map_type='DBNormalCDF', map_params=None)
    mu, sg = stats.packaged('resnet-1.1c')
    return it.set_stats(mu, sg), mu, sg


@pytest.mark.gpu
def test_mix_matches_oracle_and_rejects_bad_arguments():
    it, _, _ = _it()
    s, d, s_len, d_len, snr, off = _corpus()
    so, do, xo, nfr = it.mix(s, d, s_len, d_len, snr, off)
    r_s, r_d, r_x, r_n = train_tgt.mix(s, d, s_len, d_len, snr, off)
    assert nfr == r_n and tuple(so.shape) == r_s.shape
    assert np.array_equal(so.cpu().numpy(), r_s)
    assert np.allclose(do.cpu().numpy(), r_d, rtol=2e-6, atol=1e-9)           # alpha: one float32 sqrt / div / mean
    assert np.allclose(xo.cpu().numpy(), r_x, rtol=0, atol=2e-7)
    with pytest.raises(ValueError):
        it.mix(s, d, s_len, [l - 1 for l in s_len], snr, off)                 # noise shorter than speech
    with pytest.raises(ValueError):
        it.mix(s, d, s_len, d_len, snr, [dl - sl + 1 for sl, dl in zip(s_len, d_len)])
    _, _, _, n2 = it.mix(s, d, s_len, d_len, snr)                             # offsets drawn like sig.py:277
    assert n2 == r_n


@pytest.mark.gpu
def test_example_and_xi_match_oracle():
    it, mu, sg = _it()
    s, d, s_len, d_len, snr, off = _corpus()
    X, xi_bar, nfr = it.example(s, d, s_len, d_len, snr, off)
    r_X, r_xb, r_n = train_tgt.example(s, d, s_len, d_len, snr, off, mu, sg)
    assert nfr == r_n and tuple(X.shape) == r_X.shape
    X, xi_bar = X.cpu().numpy(), xi_bar.cpu().numpy()
    for i, n in enumerate(nfr):
        assert (np.abs(X[i, :n] - r_X[i, :n]).max(axis=-1) <= 1e-5 * r_X[i, :n].max(axis=-1)).all()
        err = np.abs(xi_bar[i, :n] - r_xb[i, :n])
        # xi is a ratio of two spectra: a bin where the noise spectrum is ~1e-4 of the frame maximum amplifies the 1e-5
        # STFT tolerance, so the bound is on the bulk and on the worst case separately
        assert np.quantile(err, 0.999) < 2e-4 and err.max() < 5e-2
        assert xi_bar[i, :n].min() >= 0.0 and xi_bar[i, :n].max() <= 1.0
    S = np.abs(np.random.default_rng(0).standard_normal((50, 257))).astype(np.float32)
    D = np.abs(np.random.default_rng(1).standard_normal((50, 257))).astype(np.float32)
    D[0, :5] = 0.0                                                             # the 1e-12 floor (sig.py:121)
    assert np.allclose(it.xi(S, D), train_tgt.xi(S, D), rtol=3e-7)
    assert np.array_equal(it.gamma(S, D), it.xi(S, D))


@pytest.mark.gpu
def test_stats_match_oracle_and_set_the_map():
    it, _, _ = _it()
    s, d, s_len, d_len, snr, off = _corpus(n=8, seed=9)
    so, do, xo, _ = train_tgt.mix(s, d, s_len, d_len, snr, off)
    acc = it.xi_db_moments(so, do, s_len).cpu().numpy()
    ref = train_tgt.xi_db_moments(so, do, s_len)
    assert np.array_equal(acc[0], ref[0])
    r_mu, r_sg = train_tgt.stats_from_moments(ref)
    mu, sg = it.stats(so, do, xo, s_len)
    assert np.abs(mu - r_mu).max() < 2e-3 and np.abs(sg - r_sg).max() < 2e-3      # dB
    assert np.array_equal(it.xi_map.mu, mu) and np.array_equal(it.xi_map.sigma, sg)
    # half + half = whole (what the all-reduce relies on)
    a = it.xi_db_moments(so[:4], do[:4], s_len[:4]) + it.xi_db_moments(so[4:], do[4:], s_len[4:])
    assert np.allclose(a.cpu().numpy(), acc, rtol=1e-12)


@pytest.mark.gpu
def test_magxigamma_and_maggain_targets_match_oracle():
    """the two other magnitude-domain targets (inp_tgt.py:345-519) on the same kernels: examples and
    enhanced speech against the oracle's primitives."""
    from deepxi_b200.inp_tgt import inp_tgt_selector
    from oracle import gain as ogain
    mu, sg = stats.packaged('resnet-1.1c')
    s, d, s_len, d_len, snr, off = _corpus(n=3, seed=13)
    so, do, xo, nfr = train_tgt.mix(s, d, s_len, d_len, snr, off)
    xg = inp_tgt_selector('MagXiGamma', 512, 256, 512, 16000, map_type=['DBNormalCDF', 'DBNormalCDF'], map_params=[None, None])
    xg.set_stats((mu, sg), (mu + 3.0, sg * 0.9))
    X, tgt, n2 = xg.example(s, d, s_len, d_len, snr, off)
    assert n2 == nfr and tuple(tgt.shape) == (3, max(nfr), 514)
    tgt = tgt.cpu().numpy()
    for i, n in enumerate(nfr):
        S, _ = sig.polar_analysis(so[i, :s_len[i]]); D, _ = sig.polar_analysis(do[i, :s_len[i]]); Xo, _ = sig.polar_analysis(xo[i, :s_len[i]])
        r_xi = cdfmap.normal_cdf_map(train_tgt.xi(S, D), mu, sg)
        r_ga = cdfmap.normal_cdf_map(train_tgt.gamma(Xo, D), mu + 3.0, sg * 0.9)
        assert np.quantile(np.abs(tgt[i, :n, :257] - r_xi), 0.999) < 2e-4 and np.quantile(np.abs(tgt[i, :n, 257:] - r_ga), 0.999) < 2e-4
    # enhanced speech from a synthetic network output: both halves inverted, gfunc, synthesis
    rng = np.random.default_rng(4)
    xb = rng.uniform(0.02, 0.98, (40, 514)).astype(np.float32)
    mag = np.abs(rng.standard_normal((40, 257))).astype(np.float32)
    pha = rng.uniform(-3.1, 3.1, (40, 257)).astype(np.float32)
    y = xg.enhanced_speech(mag, pha, xb, 'mmse-stsa')
    xi_o = cdfmap.normal_cdf_inverse(xb[:, :257], mu, sg)
    ga_o = cdfmap.normal_cdf_inverse(xb[:, 257:], mu + 3.0, sg * 0.9)
    y_ref = sig.polar_synthesis(mag * ogain.gfunc(xi_o, ga_o, 'mmse-stsa'), pha)
    err = y[:len(y_ref)] - y_ref
    assert 10 * np.log10(np.sum(y_ref.astype(np.float64) ** 2) / np.sum(err.astype(np.float64) ** 2)) > 70.0
    assert np.allclose(xg.xi_hat(xb), xi_o, rtol=3e-5) and np.allclose(xg.gamma_hat(xb), ga_o, rtol=3e-5)
    # MagGain: target = gfunc of the instantaneous SNRs; enhanced speech = |X| G_hat (thresholded for 'ibm')
    mg = inp_tgt_selector('MagGain', 512, 256, 512, 16000, gain='irm')
    X2, G, n3 = mg.example(s, d, s_len, d_len, snr, off)
    G = G.cpu().numpy()
    S, _ = sig.polar_analysis(so[0, :s_len[0]]); D, _ = sig.polar_analysis(do[0, :s_len[0]])
    r_G = ogain.gfunc(train_tgt.xi(S, D), None, 'irm')
    assert np.quantile(np.abs(G[0, :nfr[0]] - r_G), 0.999) < 2e-4 and G.min() >= 0.0 and G.max() <= 1.0
    gh = rng.uniform(0, 1, (40, 257)).astype(np.float32)
    assert np.abs(mg.enhanced_speech(mag, pha, gh)[:len(y_ref)] - sig.polar_synthesis(mag * gh, pha)).max() < 2e-6
    mi = inp_tgt_selector('MagGain', 512, 256, 512, 16000, gain='ibm')
    assert np.abs(mi.enhanced_speech(mag, pha, gh)[:len(y_ref)] - sig.polar_synthesis(mag * (gh > 0.5), pha)).max() < 2e-6
