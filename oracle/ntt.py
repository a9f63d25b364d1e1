"""Radix-2 NTT family over BLS12-381 Fr (oracle; test infrastructure).

Mirrors the API of ``poly_commit::Fft`` as the reference uses it:
``Fft::new(k)`` (``src/prover.rs:87-88``, ``src/key.rs:83,222``), ``.elements[i] = w^i``
(``src/permutation.rs:148,246``), by-value ``dft / idft / coset_dft / coset_idft`` with
zero padding of short inputs (``src/key.rs:223-245``, ``src/prover/quotient_poly.rs:54-58,
115``) and natural-order outputs, ``compute_vanishing_poly_over_coset``
(``src/key.rs:291``).  The algorithm is the textbook one the crate is believed to use
([EXT-RECALL]: bit-reverse, then radix-2 DIT butterflies; coset multiply by g^i before /
g^-i after; scale by n^-1 on the inverse).  ``dft_naive`` is the O(n^2) definition used
to check it.  Values are canonical Python ints.
"""
from .fields import R_MOD, MULTIPLICATIVE_GENERATOR, domain_generator, fr_inv

_r = R_MOD


def _bitrev_permute(a, k):
    n = 1 << k
    for i in range(n):
        j = int(format(i, "0%db" % k)[::-1], 2) if k else 0
        if i < j:
            a[i], a[j] = a[j], a[i]


def _radix2_dit(a, k, w):
    """In place: a (bit-reverse-permuted first) -> natural-order DFT with root w."""
    n = 1 << k
    _bitrev_permute(a, k)
    # twiddle table of the half-size domain
    tw = [1] * (n // 2 if n > 1 else 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * w % _r
    m = 1
    while m < n:
        stride = n // (2 * m)
        for start in range(0, n, 2 * m):
            for j in range(m):
                t = a[start + j + m] * tw[j * stride] % _r
                u = a[start + j]
                a[start + j] = (u + t) % _r
                a[start + j + m] = (u - t) % _r
        m *= 2


def dft_naive(coeffs, k, w=None):
    n = 1 << k
    w = domain_generator(k) if w is None else w
    a = list(coeffs) + [0] * (n - len(coeffs))
    out = []
    for j in range(n):
        x = pow(w, j, _r)
        acc = 0
        for c in reversed(a):
            acc = (acc * x + c) % _r
        out.append(acc)
    return out


class Fft:
    def __init__(self, k):
        self.k = k
        self.n = 1 << k
        self.w = domain_generator(k)
        self.w_inv = fr_inv(self.w)
        self.n_inv = fr_inv(self.n % _r)
        self.g = MULTIPLICATIVE_GENERATOR
        self.g_inv = fr_inv(self.g)
        self._elements = None

    # --- accessors (src/key.rs:205-207, src/prover.rs:252,300)
    def size(self):
        return self.n

    def size_inv(self):
        return self.n_inv

    def generator(self):
        return self.w

    def generator_inv(self):
        return self.w_inv

    @property
    def elements(self):
        if self._elements is None:
            e = [1] * self.n
            for i in range(1, self.n):
                e[i] = e[i - 1] * self.w % _r
            self._elements = e
        return self._elements

    def _pad(self, v):
        v = [x % _r for x in v]
        assert len(v) <= self.n, "input longer than the domain"
        return v + [0] * (self.n - len(v))

    # --- transforms
    def dft(self, coeffs):
        a = self._pad(coeffs)
        _radix2_dit(a, self.k, self.w)
        return a

    def idft(self, evals):
        a = self._pad(evals)
        _radix2_dit(a, self.k, self.w_inv)
        return [x * self.n_inv % _r for x in a]

    def coset_dft(self, coeffs):
        a = self._pad(coeffs)
        gi = 1
        for i in range(len(a)):
            a[i] = a[i] * gi % _r
            gi = gi * self.g % _r
        _radix2_dit(a, self.k, self.w)
        return a

    def coset_idft(self, evals):
        a = self.idft(evals)
        gi = 1
        for i in range(len(a)):
            a[i] = a[i] * gi % _r
            gi = gi * self.g_inv % _r
        return a

    def compute_vanishing_poly_over_coset(self, n):
        """Z_H(g * w_this^i) = (g w^i)^n - 1 for i < size (src/key.rs:291)."""
        out = []
        gn = pow(self.g, n, _r)
        wn = pow(self.w, n, _r)
        x = gn
        for _ in range(self.n):
            out.append((x - 1) % _r)
            x = x * wn % _r
        return out


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % _r
    return acc
