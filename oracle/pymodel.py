"""Independent pure-Python (big-int) model of the hot-path mathematics.  TEST INFRASTRUCTURE ONLY.

It does NOT follow the reference's code structure; it states the *mathematical* contract
 so that the C oracle (which does follow the reference line by line) can be
cross-checked against something that shares no code with it:
  A.1 minimal primitive 2N-th root, A.2 transform as direct evaluation, A.3 monomials,
  A.4 exact residues, A.5 gadget decomposition (single big integer, any width),
  A.6 external product as a schoolbook identity.
"""
from __future__ import annotations


def brv(i, bits):
    r = 0
    for b in range(bits):
        r |= ((i >> b) & 1) << (bits - 1 - b)
    return r


def min_primitive_root(log_degree, q):
    """Smallest x with x^(2^(log_degree-1)) == -1 (mod q)  <=>  primitive 2^log_degree-th root."""
    deg = 1 << log_degree
    if (q - 1) % deg:
        return None
    for r in range(2, 5000):
        w = pow(r, (q - 1) // deg, q)
        if pow(w, deg // 2, q) == q - 1:
            break
    else:
        return None
    sq, cur, best = w * w % q, w, w
    for _ in range(deg // 2):
        best = min(best, cur)
        cur = cur * sq % q
    return best


def ntt_forward(x, q, psi):
    n = len(x); logn = n.bit_length() - 1
    return [sum(int(x[j]) * pow(psi, (2 * brv(i, logn) + 1) * j, q) for j in range(n)) % q for i in range(n)]


def negacyclic_mul(a, b, q):
    n = len(a); c = [0] * n
    for i in range(n):
        for j in range(n):
            k = i + j
            if k < n:
                c[k] = (c[k] + int(a[i]) * int(b[j])) % q
            else:
                c[k - n] = (c[k - n] - int(a[i]) * int(b[j])) % q
    return c


def mul_monomial(p, r, q):
    n = len(p); out = [0] * n
    for i in range(n):
        k = i + r; sign = 1
        while k >= n:
            k -= n; sign = -sign
        out[k] = (sign * int(p[i])) % q
    return out


class Gadget:
    """Approximate signed decomposition of Z_Q (Q odd, any size) in base 2^beta."""

    def __init__(self, Q, beta, levels=None):
        self.Q, self.beta = Q, beta
        bits = Q.bit_length()
        full = bits // beta
        self.levels = levels or full
        assert 0 < self.levels <= full
        self.drop = bits - self.levels * beta
        self.B = 1 << beta
        if beta == 1:
            thr = None
            if self.drop:
                thr = ((1 << (self.levels + 1)) - 1) << (self.drop - 1)
        else:
            v = 0
            for _ in range(self.levels):
                v = (v << beta) | ((self.B - 1) >> 1)
            thr = (((v << 1) | 1) << (self.drop - 1)) if self.drop else v + 1
        self.threshold = thr if (thr is not None and thr < Q) else None
        self.add = ((1 << bits) - 1) - (Q - 1)

    def scalars(self):
        return [1 << (self.drop + l * self.beta) for l in range(self.levels)]

    def init(self, v):
        if self.threshold is not None and v >= self.threshold:
            v += self.add
        carry = (v >> (self.drop - 1)) & 1 if self.drop else 0
        return v, carry

    def unsigned_digits(self, v):
        adj, carry = self.init(int(v))
        out = []
        for l in range(self.levels):
            t = ((adj >> (self.drop + l * self.beta)) & (self.B - 1)) + carry
            carry = 1 if (t & (2 if self.beta == 1 else (self.B | (self.B >> 1)))) else 0
            out.append(t & (self.B - 1))
        return out

    def signed_digits(self, v):
        """Centred digits as integers in (-B/2, B/2] (beta > 1) -- the lift rule d < ceil(B/2) ? d : d - B."""
        half = (self.B + 1) // 2
        return [d if (self.B == 2 or d < half) else d - self.B for d in self.unsigned_digits(v)]


def crt_compose(residues, moduli):
    Q = 1
    for m in moduli:
        Q *= m
    x = 0
    for r, m in zip(residues, moduli):
        Mi = Q // m
        x = (x + (int(r) * pow(Mi, -1, m) % m) * Mi) % Q
    return x
