"""Deterministic synthetic genotype data (SURVEY.md §8d distribution).

A stateless counter-based generator: every (site, haplotype) cell is a pure
function of ``(seed, site, hap)``, so the SAME matrix can be produced
 * here in numpy (CPU tests, golden fixtures, the oracle's inputs), and
 * on the device by ``pg_synth_fill`` (csrc/pgwin.cu) for the 10M-site bench
   without writing gigabytes of text.
Both twins use only 64-bit integer mixing and integer threshold compares, so
they are bit-identical (tests/test_synth.py checks this on the GPU).

Codes: A=0 C=1 G=2 T=3, missing = -1 (the reference's ``seqNumDict``,
genomics.py:33, maps N to -999; any negative value is "missing" here).
"""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
GOLD = np.uint64(0x9E3779B97F4A7C15)
K_STREAM = np.uint64(0xD1B54A32D192ED03)
K_HAP = np.uint64(0xA24BAED4963EE407)
K_SAMPLE = np.uint64(0x9FB21C651E98DF25)

BASES = "ACGT"


def mix64(x):
    """splitmix64 finaliser on a uint64 array (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= _M1
        x ^= x >> np.uint64(27)
        x *= _M2
        x ^= x >> np.uint64(31)
    return x


def _thr(p: float) -> np.uint64:
    """Probability -> 32-bit integer threshold (event iff r32 < thr)."""
    return np.uint64(min(max(int(round(p * 4294967296.0)), 0), 4294967296))


class SynthSpec:
    """Shape + distribution parameters shared by the numpy and CUDA twins."""

    def __init__(self, n_pops, samples_per_pop, ploidy=2, seed=20260923,
                 p_variable=0.30, p_out_zero=0.80, p_third=0.01, miss=0.0):
        self.n_pops = int(n_pops)
        self.samples_per_pop = int(samples_per_pop)
        self.ploidy = int(ploidy)
        self.seed = int(seed)
        self.p_variable = float(p_variable)
        self.p_out_zero = float(p_out_zero)
        self.p_third = float(p_third)
        self.miss = float(miss)

    @property
    def n_samples(self):
        return self.n_pops * self.samples_per_pop

    @property
    def n_haps(self):
        return self.n_samples * self.ploidy

    @property
    def haps_per_pop(self):
        return self.samples_per_pop * self.ploidy

    def thresholds(self):
        return (int(_thr(self.p_variable)), int(_thr(self.p_out_zero)),
                int(_thr(self.p_third)), int(_thr(self.miss)))

    def sample_names(self):
        return ["p%d_%03d" % (p, i) for p in range(self.n_pops)
                for i in range(self.samples_per_pop)]

    def pop_names(self):
        return ["pop%d" % p for p in range(self.n_pops)]

    def hap_pop(self):
        return np.repeat(np.arange(self.n_pops, dtype=np.int32), self.haps_per_pop)


def synth_genotypes(spec: SynthSpec, site0: int, n_sites: int, chunk: int = 4096) -> np.ndarray:
    """int8 [n_sites, n_haps] for global site indices site0 .. site0+n_sites-1."""
    out = np.empty((int(n_sites), spec.n_haps), dtype=np.int8)
    for lo in range(0, int(n_sites), chunk):
        n = min(chunk, int(n_sites) - lo)
        out[lo:lo + n] = _synth_block(spec, site0 + lo, n)
    return out


def _synth_block(spec: SynthSpec, site0: int, n_sites: int) -> np.ndarray:
    S, H, P = int(n_sites), spec.n_haps, spec.n_pops
    thr_var, thr_out0, thr_third, thr_miss = (np.uint64(t) for t in spec.thresholds())
    with np.errstate(over="ignore"):
        site = np.arange(site0, site0 + S, dtype=np.uint64)
        base = mix64(np.uint64(spec.seed) * GOLD + site)            # [S]

        def draw(k):
            return mix64(base + np.uint64(k) * K_STREAM + np.uint64(1))

        d0 = draw(0)
        ref = (d0 >> np.uint64(62)).astype(np.int64)                 # 0..3
        alt = (ref + 1 + ((d0 >> np.uint64(40)) % np.uint64(3)).astype(np.int64)) % 4
        variable = (draw(1) >> np.uint64(32)) < thr_var
        out0 = (draw(2) >> np.uint64(32)) < thr_out0
        third_site = variable & ((draw(3) >> np.uint64(32)) < thr_third)
        third_pop = ((draw(4) >> np.uint64(40)) % np.uint64(P)).astype(np.int64)
        # third allele = smallest base that is neither ref nor alt
        third = np.zeros(S, dtype=np.int64)
        for cand in (3, 2, 1, 0):
            ok = (ref != cand) & (alt != cand)
            third = np.where(ok, cand, third)
        freq = np.empty((S, P), dtype=np.uint64)
        for x in range(P):
            freq[:, x] = draw(8 + x) >> np.uint64(32)
        freq[~variable, :] = 0
        freq[out0, P - 1] = 0

        hap = np.arange(H, dtype=np.uint64)
        hh = mix64(base[:, None] ^ ((hap[None, :] + np.uint64(1)) * K_HAP))   # [S,H]
        r32 = hh >> np.uint64(32)
        hp = spec.hap_pop().astype(np.int64)
        is_alt = r32 < freq[:, hp]
        g = np.where(is_alt, alt[:, None], ref[:, None])
        use_third = is_alt & third_site[:, None] & (hp[None, :] == third_pop[:, None]) \
            & (((hh >> np.uint64(8)) & np.uint64(1)) == np.uint64(1))
        g = np.where(use_third, third[:, None], g)
        if spec.miss > 0:
            samp = np.arange(spec.n_samples, dtype=np.uint64)
            mm = mix64(base[:, None] + (samp[None, :] + np.uint64(1)) * K_SAMPLE)
            miss = (mm >> np.uint64(32)) < thr_miss                           # [S,n_samples]
            g = np.where(np.repeat(miss, spec.ploidy, axis=1), -1, g)
    return g.astype(np.int8)


def synth_positions(n_sites: int, seed: int = 20260923, spacing: int = 10) -> np.ndarray:
    """Strictly increasing int32 positions, one site per ~`spacing` bp:
    pos[i] = spacing*i + 1 + (hash(i) % spacing). Same formula in pg_synth_positions."""
    with np.errstate(over="ignore"):
        i = np.arange(n_sites, dtype=np.uint64)
        h = mix64(np.uint64(seed) * GOLD + i + np.uint64(0x5851F42D4C957F2D))
        off = (h >> np.uint64(33)) % np.uint64(spacing)
    return (i * np.uint64(spacing) + np.uint64(1) + off).astype(np.int64).astype(np.int32)


def write_geno(path, geno: np.ndarray, positions, scaffolds, sample_names, ploidy=2,
               fmt="phased", sep="|"):
    """Write a .geno text file (the format parseGenoLine reads, genomics.py:1884-1904)."""
    lut = np.array(list("ACGT") + ["N"])
    S, H = geno.shape
    assert H == len(sample_names) * ploidy
    codes = np.where(geno < 0, 4, geno).astype(np.int64)
    chars = lut[codes]                                   # [S,H] of 1-char str
    diplo = {"AA": "A", "CC": "C", "GG": "G", "TT": "T", "GT": "K", "TG": "K", "AC": "M", "CA": "M",
             "CG": "S", "GC": "S", "AG": "R", "GA": "R", "AT": "W", "TA": "W", "CT": "Y", "TC": "Y"}
    with open(path, "wt") as f:
        f.write("#CHROM\tPOS\t" + "\t".join(sample_names) + "\n")
        for s in range(S):
            row = chars[s]
            if fmt == "phased":
                toks = [sep.join(row[i * ploidy:(i + 1) * ploidy]) for i in range(len(sample_names))]
            elif fmt == "haplo":
                toks = list(row)
            elif fmt == "pairs":
                toks = ["".join(row[i * 2:i * 2 + 2]) for i in range(len(sample_names))]
            elif fmt == "diplo":
                toks = [diplo.get(row[2 * i] + row[2 * i + 1], "N") for i in range(len(sample_names))]
            else:
                raise ValueError(fmt)
            f.write("%s\t%d\t%s\n" % (scaffolds[s], int(positions[s]), "\t".join(toks)))
