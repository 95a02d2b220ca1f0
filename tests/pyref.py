"""Independent big-int restatement of the reference algorithm (pure Python, small sizes only).

Second implementation used to cross-check the C++ oracle bit-for-bit.  Written
from the algorithm statement (reference: src/argument_witness_calc.rs:87-136,
src/regular_functions_utils.rs:266-480, src/negbase_utils.rs:20-36), with schoolbook polynomial
products and affine curve arithmetic on Python ints -- it shares no code with oracle/.
Test infrastructure only.
"""
from math import isqrt

FIELDS = {
    "pallas_fp": 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001,
    "pallas_fq": 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001,
    "bn256_fr": 21888242871839275222246405745257275088548364400416034343698204186575808495617,
    "bn256_fq": 21888242871839275222246405745257275088696311157297823662689037894645226208583,
}
FIELD_ID = {"pallas_fp": 0, "pallas_fq": 1, "bn256_fr": 2, "bn256_fq": 3}
# name: (id, base field, scalar field, b)
CURVES = {
    "pallas": (0, "pallas_fp", "pallas_fq", 5),
    "vesta": (1, "pallas_fq", "pallas_fp", 5),
    "grumpkin": (2, "bn256_fr", "bn256_fq", -17),
}
R = 1 << 256


class Curve:
    def __init__(self, name):
        self.name = name
        self.id, bf, sf, self.b = CURVES[name]
        self.p = FIELDS[bf]
        self.q = FIELDS[sf]
        self.base_field, self.scalar_field = bf, sf

    # affine points are (x, y) tuples, identity is None
    def neg(self, P):
        return None if P is None else (P[0], (-P[1]) % self.p)

    def add(self, P, Q):
        p = self.p
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if (P[1] + Q[1]) % p == 0:
                return None
            lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, p) % p
        else:
            lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
        x = (lam * lam - P[0] - Q[0]) % p
        return (x, (lam * (P[0] - x) - P[1]) % p)

    def mul(self, k, P):
        acc = None
        for bit in bin(k)[2:] if k else "":
            acc = self.add(acc, acc)
            if bit == "1":
                acc = self.add(acc, P)
        return acc

    def on_curve(self, P):
        return P is None or (P[1] * P[1] - P[0] ** 3 - self.b) % self.p == 0

    def sqrt(self, a):
        """Tonelli-Shanks in the base field; None for non-residues."""
        p = self.p
        a %= p
        if a == 0:
            return 0
        if pow(a, (p - 1) // 2, p) != 1:
            return None
        s, t = 0, p - 1
        while t % 2 == 0:
            s, t = s + 1, t // 2
        z = 2
        while pow(z, (p - 1) // 2, p) != p - 1:
            z += 1
        m, c, tt, r = s, pow(z, t, p), pow(a, t, p), pow(a, (t + 1) // 2, p)
        while tt != 1:
            i, x = 0, tt
            while x != 1:
                x, i = x * x % p, i + 1
            bb = pow(c, 1 << (m - i - 1), p)
            m, c = i, bb * bb % p
            tt, r = tt * c % p, r * bb % p
        return r


# multiplicative generators ff's field impls derive ROOT_OF_UNITY = g^t from (p - 1 = 2^S t)
FIELD_GENERATOR = {"pallas_fp": 5, "pallas_fq": 5, "bn256_fr": 7}


def sqrt_ff(field, a):
    """ff::helpers::sqrt_tonelli_shanks restated: the root a^((t+1)/2) * z^e with z = ROOT_OF_UNITY = g^t.
    Returns (is_square, root); for a non-residue the root is sqrt(ROOT_OF_UNITY * a) like Field::sqrt_alt."""
    p = FIELDS[field]
    s, t = 0, p - 1
    while t % 2 == 0:
        s, t = s + 1, t // 2
    root = pow(FIELD_GENERATOR[field], t, p)

    def ts(v):
        if v % p == 0:
            return 0
        w = pow(v, (t - 1) // 2, p)
        x, b, z, m = v * w % p, v * w * w % p, root, s
        while b != 1:
            k, b2 = 0, b
            while b2 != 1:
                b2, k = b2 * b2 % p, k + 1
                if k == m:
                    return None
            zz = pow(z, 1 << (m - k - 1), p)
            x, z = x * zz % p, zz * zz % p
            b, m = b * z % p, k
        return x
    r = ts(a % p)
    if r is not None:
        return True, r
    return False, ts(a * root % p)


def negbase_decompose(x, base):
    acc = []
    while x != 0:
        digit = x % base  # Python's % is already non-negative
        acc.append(digit)
        x = -((x - digit) // base)
    return acc


def logb_ceil(x, base):
    i = 0
    while x > 0:
        x //= base
        i += 1
    return i


def num_digits(curve, base):
    return logb_ceil(isqrt(curve.q) + 2, base) + 1


# ---- polynomials: python lists of ints mod p, low degree first, reference lengths kept -----------
def pmul(a, b, p):
    if len(a) + len(b) == 0:
        return []
    r = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                r[i + j] = (r[i + j] + x * y) % p
    return r


def padd(a, b, p):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % p for i in range(n)]


def kate_div(a, root, p):
    q = [0] * (len(a) - 1)
    tmp = 0
    for i in range(len(a) - 2, -1, -1):
        q[i] = (a[i + 1] + tmp) % p
        tmp = q[i] * root % p
    return q


def peval(a, x, p):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % p
    return acc


def rf_mul(f, g, cv):
    p = cv.p
    subst = [cv.b % p, 0, 0, 1]
    return (padd(pmul(f[0], g[0], p), pmul(pmul(f[1], g[1], p), subst, p), p),
            padd(pmul(f[0], g[1], p), pmul(f[1], g[0], p), p))


def proj(P):
    return (0, 0, 0) if P is None else (P[0], P[1], 1)


def linefunc(A, B, cv):
    p = cv.p
    ax, ay, az = proj(A)
    bx, by, bz = proj(B)
    lz, lx, ly = (ax * by - ay * bx) % p, (ay * bz - az * by) % p, (az * bx - ax * bz) % p
    if lx or ly or lz:
        return ([lz, lx], [ly])
    cx, cy, cz = proj(cv.neg(cv.add(A, B)))
    return ([(ax * cy - ay * cx) % p, (ay * cz - az * cy) % p], [(az * cx - ax * cz) % p])


def from_point(P, cv):
    if P is None:
        return (None, ([1], []))
    return (cv.neg(P), linefunc(P, cv.neg(P), cv))


def from_pair(P, Q, cv):
    if P is None:
        return from_point(Q, cv)
    return (cv.neg(cv.add(P, Q)), linefunc(P, Q, cv))


def merge(a, b, cv):
    p = cv.p
    out = cv.add(a[0], b[0])
    if a[0] is None or b[0] is None:
        return (out, rf_mul(a[1], b[1], cv))
    num = rf_mul(a[1], rf_mul(b[1], linefunc(cv.neg(a[0]), cv.neg(b[0]), cv), cv), cv)
    ax, bx = a[0][0], b[0][0]
    return (out, (kate_div(kate_div(num[0], ax, p), bx, p), kate_div(kate_div(num[1], ax, p), bx, p)))


def divisor_witness_partial(pts, cv):
    if not pts:
        return (([1], []), None)
    level = []
    i = 0
    while i < len(pts) - 1:
        level.append(from_pair(pts[i], pts[i + 1], cv))
        i += 2
    if i == len(pts) - 1:
        level.append(from_point(pts[i], cv))
    while len(level) > 1:
        nxt = []
        for k in range(0, len(level), 2):
            nxt.append(merge(level[k], level[k + 1], cv) if k + 1 < len(level) else level[k])
        level = nxt
    return (level[0][1], level[0][0])


def divisor_witness(pts, cv):
    f, out = divisor_witness_partial(pts, cv)
    assert out is None, "points do not sum to identity"
    return f


def divisor_witness_naive(pts, cv):
    """reference: src/regular_functions_utils.rs:483-551; lines as (lx, ly, lz)"""
    pos, neg, rpos, rneg, tmp = list(pts), [], [], [], []

    def drain(lst):
        while len(lst) > 1:
            inc1 = lst.pop()
            if inc1 is not None:
                tmp.append((inc1, lst.pop()))

    def flush(lines, sums):
        out = []
        for a, b in tmp:
            la, lb = linefunc(a, b, cv)
            out.append(((la[1], lb[0], la[0]), cv.neg(cv.add(a, b))))
        tmp.clear()
        while out:
            line, s = out.pop()
            lines.append(line)
            sums.append(s)

    while len(pos) > 1 or len(neg) > 1:
        drain(pos)
        flush(rpos, neg)
        drain(neg)
        flush(rneg, pos)
    ok = (not pos and not neg) or (len(pos) == 1 and not neg and pos[0] is None) or (not pos and len(neg) == 1 and neg[0] is None) \
        or (len(pos) == 1 and len(neg) == 1 and pos[0] == neg[0])
    assert ok, "points do not sum to identity"
    return rpos, rneg


def _wrap_i128(v):
    v &= (1 << 128) - 1
    return v - (1 << 128) if v >> 127 else v


def prepare_scalar_witness(sc, base, num_digits, logtable, intended=False):
    """reference: src/negbase_utils.rs:79-124.  Returns rows[base][num_limbs+1] of ("scalar", sc) / ("bucket", v) / ("limb", v, mask);
    i128 sums wrap (release-build semantics); intended=True uses limb slot i // logtable + 1 instead of i % logtable + 1"""
    digits = negbase_decompose(sc, base)
    assert len(digits) <= num_digits
    num_limbs = (num_digits + logtable - 1) // logtable
    ret = [[[0, 0] for _ in range(num_limbs + 1)] for _ in range(base)]
    for i, dg in enumerate(digits):
        if dg == 0:
            continue
        slot = (i // logtable if intended else i % logtable) + 1
        e = i % logtable
        if slot > num_limbs:
            raise IndexError("limb slot out of bounds")
        ret[dg][0][0] += (-base) ** i
        for row in (dg, 0):
            ret[row][slot][0] += (-base) ** e
            ret[row][slot][1] += 2 ** e
    out = []
    for i in range(base):
        row = []
        for j in range(num_limbs + 1):
            if i == 0 and j == 0:
                row.append(("scalar", sc))
            elif j == 0:
                row.append(("bucket", _wrap_i128(ret[i][j][0])))
            else:
                row.append(("limb", _wrap_i128(ret[i][j][0]), ret[i][j][1] & 0xFFFFFFFF))
        out.append(row)
    return out


def lhs_witness(scalars, pts, base, cv):
    """scalars: canonical ints; pts: affine tuples / None.  Returns (digits, carries, fns)."""
    assert len(scalars) == len(pts)
    sq = isqrt(cv.q) + 2
    d = logb_ceil(sq, base) + 1
    digits = []
    for s in scalars:
        assert s < sq
        dg = negbase_decompose(s, base)
        assert len(dg) <= d
        dg = dg + [0] * (d - len(dg))
        digits.append(dg[::-1])
    mult = []
    for P in pts:
        acc, row = P, []
        for _ in range(1, base):
            row.append(acc)
            acc = cv.add(acc, P)
        mult.append(row)
    carry, carries, ret = None, [], []
    for i in range(d):
        tmp = []
        if carry is not None:
            tmp += [cv.neg(carry)] * base
        carry = cv.mul(base, cv.neg(carry))
        for j in range(len(pts)):
            dg = digits[j][i]
            if dg:
                tmp.append(mult[j][dg - 1])
                carry = cv.add(carry, mult[j][dg - 1])
        tmp.append(cv.neg(carry))
        carries.append(carry)
        ret.append(divisor_witness(tmp, cv))
    ret.reverse()
    return digits, carries, ret


def canonicalize(f, p):
    a, b = list(f[0]), list(f[1])
    while a and a[-1] == 0:
        a.pop()
    while b and b[-1] == 0:
        b.pop()
    if not a and not b:
        return ([], [])
    oa = 2 * (len(a) - 1) if a else -1
    ob = 2 * (len(b) - 1) + 3 if b else -1
    lead = a[-1] if oa > ob else b[-1]
    inv = pow(lead, -1, p)
    return ([c * inv % p for c in a], [c * inv % p for c in b])


def rf_eval(f, P, p):
    return (peval(f[0], P[0], p) + peval(f[1], P[0], p) * P[1]) % p


# ---- deterministic synthetic inputs: SplitMix64 ---------------------------------
class SplitMix64:
    def __init__(self, seed):
        self.s = seed & (R - 1) & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def next_bits(self, nwords):
        v = 0
        for i in range(nwords):
            v |= self.next() << (64 * i)
        return v


def random_point(rng, cv):
    """try-and-increment: x from the PRNG, y = sqrt(x^3 + b), even y."""
    x = rng.next_bits(4) % cv.p
    while True:
        y = cv.sqrt(x ** 3 + cv.b)
        if y is not None and y != 0:
            if y & 1:
                y = cv.p - y
            return (x, y)
        x = (x + 1) % cv.p


def random_scalar(rng, cv):
    return rng.next_bits(2) % (isqrt(cv.q) + 2)


# ---- Montgomery packing helpers shared by the ctypes wrappers -------------------------------------
def to_mont_words(v, p):
    m = v * R % p
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_mont_words(w, p):
    m = sum(int(w[i]) << (64 * i) for i in range(4))
    return m * pow(R, -1, p) % p
