"""CPU oracle of the gradient post-processing (TEST INFRASTRUCTURE ONLY -- never imported by the product).

Independent numpy statement of what ADFWI/propagator/gradient_process.py computes (reference file:line in the
comments), float64 wherever the reference promotes to float64 and float32 where its in-place products stay float32.
It is written from the maths, not from the reference's code: the Gaussian filter is built as an outer product of its
1-D factor (the reference's covariance is diagonal, gradient_process.py:40, so its 2-D filter factorises exactly up
to rounding), while the convolution itself stays the direct 2-D sum of `scipy.signal.convolve2d`, which is the
third-party routine the reference calls (:45-46).

Pinned by tests/golden/gradproc_*.npz (outputs of the unmodified reference, tests/golden/make_golden_gradproc.py) to
1e-12 relative: tests/test_gradproc_oracle.py.
"""
import numpy as np
from scipy.signal import convolve2d


def gauss_factor(span):
    """1-D factor f of the smoothing filter: taps at -2*span .. 2*span in steps of 2 (2*span+1 of them, :36-37),
    standard deviation `span` (:40), unit sum (:42 -- the 1/(2*pi*sqrt(D)) of :25 cancels)."""
    t = np.linspace(-2.0 * span, 2.0 * span, 2 * span + 1)
    e = np.exp(-0.5 * (t / float(span)) ** 2)
    return e / e.sum()


def smooth2d(plane, span=10):
    """Normalised zero-boundary smoothing (:30-49): (plane * F) / (1 * F), F = f f^T, 'same' extent."""
    f = gauss_factor(span)
    F = np.outer(f, f)
    num = convolve2d(np.array(plane), F, mode="same")
    den = convolve2d(np.ones(np.shape(plane)), F, mode="same")
    return num / den


def taper_plane(nz, nx, size, thred, marine):
    """Mute / damping weights (:51-72).  Marine: zero on the first `size` rows.  Land: the falling half of a Hamming
    window of length 2*size laid along the first `size` COLUMNS of every row (sic, :63-64), smoothed with span size//2,
    scaled to a maximum of 1-thred, flipped (1 - t) and squared."""
    if marine:
        w = np.ones((nz, nx))
        w[:size] = 0.0
        return w
    half = np.hamming(2 * size)[size:]          # scipy.signal.hamming of the reference == numpy.hamming
    w = np.zeros((nz, nx))
    w[:, :size] = half[None, :]
    w = smooth2d(w, span=size // 2)
    w = w / w.max()
    w = 1.0 - w * (1.0 - thred)
    return w * w


def illumination_span(nz, nx):
    """Span of the illumination smoothing (:112-115)."""
    m = min(nz, nx)
    return 40 if m > 40 else int(m / 2)


def grad_process(nx, nz, vmax, grad, forw=None, grad_mute=0, grad_smooth=0, grad_mask=None, norm_grad=True,
                 forw_illumination=True, marine_or_land="land"):
    """GradProcessor.forward (:88-135) on a float32 (or float64) gradient plane; returns a new array."""
    kind = marine_or_land.lower()
    if kind not in ("marine", "offshore", "land", "onshore"):
        raise ValueError("not supported modeling marine_or_land: %s" % marine_or_land)
    thred = 0.0 if kind in ("marine", "offshore") else 0.001                       # :90-95
    g = np.array(grad)                                                             # dtype of the caller's plane
    if grad_mute > 0:                                                              # :97-98, product stored in g's dtype
        # grad_taper compares against the capitalised names (:55): anything else takes the land branch
        g *= taper_plane(nz, nx, grad_mute, thred, marine_or_land in ("Marine", "Offshore"))
    if grad_mask is not None:                                                      # :101-107
        if np.shape(grad_mask) != np.shape(g):
            raise ValueError("Wrong size of grad mask")
        g *= grad_mask
    if forw_illumination and forw is not None:                                     # :117-122
        p = smooth2d(forw, illumination_span(nz, nx))
        p = p / np.max(p + 1e-5)
        p = np.maximum(p, 0.0001)
        g = g / (p * p)
    if grad_smooth > 0:                                                            # :125-131 (case-sensitive test, sic)
        if marine_or_land in ("marine", "offshore"):
            g[grad_mute:] = smooth2d(g[grad_mute:], span=grad_smooth)
        else:
            g = smooth2d(g, span=grad_smooth)
    if norm_grad:                                                                  # :134-135
        g = vmax * g / np.abs(g).max()
    return g
