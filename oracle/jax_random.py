"""TEST INFRASTRUCTURE (oracle): numpy restatement of the parts of jax.random the reference's hot path calls
(fee_jax.py:186,237-255,271; detsim_jax.py:393; sim_jax.py:359-360,757): typed threefry keys, split, normal.

jax / jaxlib are third-party and absent from this image; the algorithm restated here is the published one:
Threefry-2x32 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; 20 rounds, rotations
13,15,26,6 / 17,29,16,24, key-schedule parity 0x1BD11BDA) driven the way jax/_src/prng.py drives it.  Pinned by
(i) the Random123 known-answer vectors that JAX's own test-suite checks (tests/random_test.py::testThreefry2x32) and
(ii) the values printed in the JAX documentation for random.normal(random.key(42)) in both counter layouts
(tests/test_oracle_golden.py).  Two counter layouts exist: ``partitionable=True`` (jax_threefry_partitionable, the
default since JAX 0.5.0, which the reference's CI uses) and the original one.
erfinv follows XLA's float32 ErfInv (Giles' single-precision polynomial); XLA's own log/sqrt may differ from numpy's
by an ulp, so normals are expected to agree with real JAX to ~1e-6 relative, the underlying random BITS exactly."""
import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, d):
    return ((x << np.uint32(d)) | (x >> np.uint32(32 - d))).astype(np.uint32)


def threefry2x32(k0, k1, x0, x1):
    """Elementwise Threefry-2x32-20 on uint32 arrays (keys broadcast)."""
    with np.errstate(over="ignore"):
        k0 = np.uint32(k0); k1 = np.uint32(k1)
        x0 = np.asarray(x0, dtype=np.uint32).copy(); x1 = np.asarray(x1, dtype=np.uint32).copy()
        ks = (k0, k1, np.uint32(k0 ^ k1 ^ np.uint32(0x1BD11BDA)))
        x0 = (x0 + ks[0]).astype(np.uint32); x1 = (x1 + ks[1]).astype(np.uint32)
        for r in range(5):
            for d in _ROT[r % 2]:
                x0 = (x0 + x1).astype(np.uint32)
                x1 = _rotl(x1, d) ^ x0
            x0 = (x0 + ks[(r + 1) % 3]).astype(np.uint32)
            x1 = (x1 + ks[(r + 2) % 3] + np.uint32(r + 1)).astype(np.uint32)
    return x0, x1


def key(seed):
    """jax.random.key(seed) / PRNGKey(seed): key data (hi, lo) of the 64-bit seed."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return (np.uint32(seed >> 32), np.uint32(seed & 0xFFFFFFFF))


def _bits_original(k, n):
    """threefry_2x32(key, iota(n)): the count array is cut into two halves that form the two words of each block."""
    cnt = np.arange(n, dtype=np.uint32)
    odd = n % 2
    if odd:
        cnt = np.concatenate([cnt, np.zeros(1, np.uint32)])
    half = cnt.size // 2
    o0, o1 = threefry2x32(k[0], k[1], cnt[:half], cnt[half:])
    out = np.concatenate([o0, o1])
    return out[:n]


def split(k, num=2, partitionable=True):
    """random.split(key, num) -> list of keys."""
    if partitionable:
        b1, b2 = threefry2x32(k[0], k[1], np.zeros(num, np.uint32), np.arange(num, dtype=np.uint32))
        return [(b1[i], b2[i]) for i in range(num)]
    flat = _bits_original(k, 2 * num)
    return [(flat[2 * i], flat[2 * i + 1]) for i in range(num)]


def random_bits(k, n, partitionable=True):
    """32 random bits per element for a total of n elements (row-major over the requested shape)."""
    if partitionable:
        idx = np.arange(n, dtype=np.uint64)
        b1, b2 = threefry2x32(k[0], k[1], (idx >> np.uint64(32)).astype(np.uint32), (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32))
        return b1 ^ b2
    return _bits_original(k, n)


def erfinv_f32(x):
    """XLA's float32 ErfInv: Giles' polynomial in w = -log((1-x)(1+x))."""
    x = np.asarray(x, dtype=np.float32)
    w = (-np.log1p((-x * x).astype(np.float32))).astype(np.float32)
    small = w < np.float32(5.0)
    ws = (w - np.float32(2.5)).astype(np.float32)
    wl = (np.sqrt(np.maximum(w, np.float32(5.0))).astype(np.float32) - np.float32(3.0)).astype(np.float32)
    cs = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503, -0.00417768164,
          0.246640727, 1.50140941)
    cl = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613, 0.00943887047,
          1.00167406, 2.83297682)
    ps = np.full_like(x, np.float32(cs[0]))
    for c in cs[1:]:
        ps = (np.float32(c) + ps * ws).astype(np.float32)
    pl = np.full_like(x, np.float32(cl[0]))
    for c in cl[1:]:
        pl = (np.float32(c) + pl * wl).astype(np.float32)
    out = (np.where(small, ps, pl) * x).astype(np.float32)
    return np.where(np.abs(x) == 1, np.float32(np.inf) * x, out).astype(np.float32)


def uniform_pm1(bits):
    """random.uniform(key, shape, float32, minval=nextafter(-1,0), maxval=1) from the raw bits."""
    fb = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).astype(np.uint32)
    fl = (fb.view(np.float32) - np.float32(1.0)).astype(np.float32)
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    hi = np.float32(1.0)
    return np.maximum(lo, (fl * (hi - lo) + lo).astype(np.float32)).astype(np.float32)


def normal(k, shape, partitionable=True):
    """random.normal(key, shape, float32) = sqrt(2) * erfinv(uniform(-1, 1))."""
    n = int(np.prod(shape)) if len(shape) else 1
    u = uniform_pm1(random_bits(k, n, partitionable))
    return (np.float32(np.sqrt(2)) * erfinv_f32(u)).astype(np.float32).reshape(shape)


def fee_noise(seed_or_key, npix, n_adc=10, partitionable=True):
    """The standard normals get_adc_values draws (fee_jax.py:186,237-255,271), laid out as the FEE kernel's noise buffer
    [base(npix) | extra(n_adc,npix) | pass(n_adc,npix) | fail(n_adc,npix)]."""
    k = key(seed_or_key) if np.isscalar(seed_or_key) else seed_or_key
    base = normal(k, (npix,), partitionable)
    kk = split(k, 1, partitionable)[0]
    extra, qpass, qfail = [], [], []
    for _ in range(n_adc):
        kk = split(kk, 1, partitionable)[0]
        extra.append(normal(kk, (npix,), partitionable))
        kk = split(kk, 1, partitionable)[0]
        qpass.append(normal(kk, (npix,), partitionable))
        kk = split(kk, 1, partitionable)[0]
        qfail.append(normal(kk, (npix,), partitionable))
    return np.concatenate([base] + extra + qpass + qfail).astype(np.float32)
