"""Turbulence / wind profile generators used to build config inputs; same names and call
signatures as the reference's `fast.turbulence_models` (fast/turbulence_models.py:4-105).
Host-side numpy: these run once per configuration in microseconds."""
import numpy


def HV57(h, w=21, A=1.7e-14):
    """Hufnagel-Valley 5/7 Cn2(h) [m^-2/3] (fast/turbulence_models.py:4-19)."""
    h = numpy.asarray(h, dtype=float)
    strato = 0.00594 * (w / 27) ** 2 * (1e-5 * h) ** 10 * numpy.exp(-h / 1000)
    tropo = 2.7e-16 * numpy.exp(-h / 1500)
    ground = A * numpy.exp(-h / 100.)
    return strato + tropo + ground


def Bufton_wind(h, vg=8, vt=30, ht=9400., Lt=4800.):
    """Bufton wind speed profile [m/s] (fast/turbulence_models.py:22-38)."""
    h = numpy.asarray(h, dtype=float)
    return vg + vt * numpy.exp(-((h - ht) / Lt) ** 2)


def equivalent_layers(h, p, L, w=None):
    """Equivalent-layers profile compression (Fusco 1999; fast/turbulence_models.py:65-105):
    split into L equal-height slabs, sum Cn2 per slab, and place each slab at its 5/3-moment
    effective height (and wind speed, conserving theta0 and tau0)."""
    edges = numpy.arange(h.min(), h.max(), (h.max() - h.min()) / L)
    which = numpy.digitize(h, edges)
    h_out = numpy.zeros(L)
    cn2_out = numpy.zeros(L)
    w_out = numpy.zeros(L) if w is not None else None
    for i in range(L):
        sel = which == i + 1
        total = p[sel].sum()
        cn2_out[i] = total
        h_out[i] = ((p[sel] * h[sel] ** (5 / 3)).sum() / total) ** (3 / 5)
        if w is not None:
            w_out[i] = ((p[sel] * w[sel] ** (5 / 3)).sum() / total) ** (3 / 5)
    if w is not None:
        return h_out, cn2_out, w_out
    return h_out, cn2_out


def HV57_Bufton_profile(N, w=21, A=1.7e-14, vg=8, vt=30, ht=9400., Lt=4800.):
    """N-layer profile: HV57 Cn2 and Bufton wind on 1 m bins up to 30 km, compressed with
    equivalent_layers (fast/turbulence_models.py:41-62).  Returns (h, cn2dh, wind)."""
    h0 = numpy.arange(0, 30000)
    return equivalent_layers(h0, HV57(h0, w, A), N, w=Bufton_wind(h0, vg, vt, ht, Lt))
