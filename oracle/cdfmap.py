"""Oracle: DBNormalCDF map (test infrastructure; see oracle/__init__.py).

Restates deepxi/map.py:62-85 (Map.db / db_inverse) and :352-402 (NormalCDF.map / inverse / stats)
in numpy float32, same operation order as the TF ops the reference issues.
"""
import numpy as np
from scipy import special as spsp

f32 = np.float32


def db(x):
    """Map.db (deepxi/map.py:62-73): 10*log(max(x,1e-12))/log(10) in f32."""
    x = np.maximum(np.asarray(x, f32), f32(1e-12))
    return f32(10.0) * (np.log(x) / np.log(f32(10.0)))


def db_inverse(x_db):
    """Map.db_inverse (deepxi/map.py:75-85): 10^(x_db/10) in f32."""
    with np.errstate(over='ignore'):
        return np.power(f32(10.0), np.asarray(x_db, f32) / f32(10.0)).astype(f32)


def normal_cdf_map(xi, mu, sigma):
    """NormalCDF.map with map_type 'DBNormalCDF' (deepxi/map.py:356-371)."""
    x = db(xi)
    v_1 = x - np.asarray(mu, f32)
    v_2 = np.asarray(sigma, f32) * np.sqrt(f32(2.0))
    v_3 = spsp.erf((v_1 / v_2).astype(f32)).astype(f32)
    return (f32(0.5) * (f32(1.0) + v_3)).astype(f32)


def normal_cdf_inverse(x_bar, mu, sigma, dtype=np.float32):
    """NormalCDF.inverse with map_type 'DBNormalCDF' (deepxi/map.py:373-390).

    xi = 10^((sigma*sqrt(2)*erfinv(2*x_bar-1) + mu)/10).  With dtype=float64 this is the
    high-precision variant used for error budgets.
    """
    t = dtype
    x_bar = np.asarray(x_bar, t)
    v_1 = np.asarray(sigma, t) * np.sqrt(t(2.0))
    v_2 = t(2.0) * x_bar
    with np.errstate(all='ignore'):
        # tf.math.erfinv in float32: restated as the correctly rounded value (float64 erfinv rounded to
        # float32); scipy's own float32 loop differs from it in the last ulp for ~1e-3 of the inputs
        v_3 = spsp.erfinv((v_2 - t(1.0)).astype(np.float64)).astype(t)
        v_4 = v_1 * v_3
        x = v_4 + np.asarray(mu, t)
        return np.power(t(10.0), x / t(10.0)).astype(t)


def normal_cdf_inverse_db(x_bar, mu, sigma, dtype=np.float64):
    """xi_hat in dB (the quantity the 0.1 dB tolerance is stated on)."""
    t = dtype
    with np.errstate(all='ignore'):
        return (np.asarray(sigma, t) * np.sqrt(t(2.0)) * spsp.erfinv(t(2.0) * np.asarray(x_bar, t) - t(1.0))
                + np.asarray(mu, t))


def normal_cdf_stats(xi):
    """NormalCDF.stats (deepxi/map.py:392-402): per-bin mean / population std of db(xi)."""
    x = db(xi)
    return x.mean(axis=0, dtype=f32), x.std(axis=0, dtype=f32)


def ibm_threshold(mu, sigma):
    """x_bar value where xi_hat crosses 1 (0 dB) in exact arithmetic: Phi(-mu/sigma)."""
    return 0.5 * (1.0 + spsp.erf(-np.asarray(mu, np.float64) / (np.asarray(sigma, np.float64) * np.sqrt(2.0))))
