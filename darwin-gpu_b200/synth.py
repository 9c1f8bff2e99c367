"""Seeded synthetic workloads for the GACT path (numpy only; no alignment logic here).

``BASELINE.json`` configs reproduced:
  * config 2 -- tile microbatch: ``tile_microbatch()`` -- reference windows of a random
    genome and query windows of the same region pushed through a PacBio-like error channel
    (15 %: 1.5 % substitutions, 9 % insertions, 4.5 % deletions),
    82 % full tiles / 18 % edge tiles, 5.5 % first tiles, direction 50/50.
  * configs 1/3/5 -- reads: ``random_genome()``, ``sample_reads()``, ``write_fasta()``
    (70-column FASTA, the only width the reference's reader accepts, fasta.h:19).
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for a, b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[a] = b


def random_genome(n, rng):
    return ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def revcomp(seq):
    """Reverse complement as the reference does it (darwin.cpp:110-147: ACGT/acgt/N)."""
    return _COMP[np.asarray(seq, dtype=np.uint8)[::-1]]


def error_channel(ref, rng, sub=0.015, ins=0.09, dele=0.045):
    """Push ``ref`` (uint8 ASCII) through a substitution/insertion/deletion channel.
    Returns (query, start) with start[i] = index in ``query`` where ref base i landed
    (start has len(ref)+1 entries; the last one is len(query))."""
    ref = np.asarray(ref, dtype=np.uint8)
    n = len(ref)
    u = rng.random(n)
    is_del = u < dele
    is_sub = (u >= dele) & (u < dele + sub)
    is_ins = (u >= dele + sub) & (u < dele + sub + ins)
    count = np.where(is_del, 0, 1) + is_ins
    base = ref.copy()
    ns = int(is_sub.sum())
    if ns:
        code = np.searchsorted(ACGT, base[is_sub])
        base[is_sub] = ACGT[(code + rng.integers(1, 4, size=ns)) % 4]
    q = np.repeat(base, count)
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(count, out=start[1:])
    ipos = start[:-1][is_ins] + 1
    q[ipos] = ACGT[rng.integers(0, 4, size=len(ipos))]
    return q, start


def tile_microbatch(n_tiles, tile_size=320, seed=42, genome_len=None, full_frac=0.82,
                    first_frac=0.055, err=(0.015, 0.09, 0.045)):
    """Config 2.  Returns dict(ref=uint8[], query=uint8[], ref_off, query_off, ref_len,
    query_len, reverse, first) with offsets into the two flat buffers.

    A tile with reverse=0 is a left-extension tile: both windows END at corresponding
    positions (the DP corner is the anchor, gact.cpp:87-94); reverse=1 is a right-extension
    tile: both windows START at corresponding positions (gact.cpp:149-156)."""
    rng = np.random.default_rng(seed)
    T = tile_size
    if genome_len is None:
        genome_len = int(min(max(n_tiles * 40, 4 * T + 1024), 64 << 20))
    ref = random_genome(genome_len, rng)
    query, start = error_channel(ref, rng, *err)
    full = rng.random(n_tiles) < full_frac
    rl = np.where(full, T, rng.integers(1, T + 1, size=n_tiles)).astype(np.int32)
    ql = np.where(full, T, rng.integers(1, T + 1, size=n_tiles)).astype(np.int32)
    reverse = (rng.random(n_tiles) < 0.5).astype(np.uint8)
    first = (rng.random(n_tiles) < first_frac).astype(np.uint8)
    # anchor in the reference, far enough from both ends for either direction
    lo, hi = 2 * T, genome_len - 2 * T
    p = rng.integers(lo, hi, size=n_tiles)
    qp = start[p]
    ref_off = np.where(reverse == 1, p, p - rl).astype(np.int64)
    query_off = np.where(reverse == 1, qp, qp - ql).astype(np.int64)
    # keep query windows inside the buffer
    query_off = np.clip(query_off, 0, len(query) - T - 1)
    return dict(ref=ref, query=query, ref_off=ref_off, query_off=query_off, ref_len=rl,
                query_len=ql, reverse=reverse, first=first)


def sample_reads(genome_seqs, n_bases, rng, mean=10000.0, sd=3000.0, lo=1000, hi=30000,
                 err=(0.015, 0.09, 0.045), revcomp_frac=0.5, prefix="S"):
    """PBSIM-like reads: log-normal lengths, sampled uniformly from the genome sequences,
    pushed through the error channel, half of them reverse-complemented.
    Returns (names, reads) with names ``<prefix><i>_<chr>_<pos>_<len>``."""
    mu = np.log(mean * mean / np.sqrt(sd * sd + mean * mean))
    sigma = np.sqrt(np.log(1.0 + (sd * sd) / (mean * mean)))
    lens_g = np.array([len(g) for g in genome_seqs], dtype=np.float64)
    names, reads, total, i = [], [], 0, 0
    while total < n_bases:
        L = int(np.clip(rng.lognormal(mu, sigma), lo, hi))
        c = int(rng.choice(len(genome_seqs), p=lens_g / lens_g.sum()))
        g = genome_seqs[c]
        if L >= len(g):
            L = len(g) - 1
        pos = int(rng.integers(0, len(g) - L))
        q, _ = error_channel(g[pos:pos + L], rng, *err)
        if rng.random() < revcomp_frac:
            q = revcomp(q)
        if len(q) == 0:
            continue
        names.append(f"{prefix}{i}_{c}_{pos}_{L}")
        reads.append(q)
        total += len(q)
        i += 1
    return names, reads


def write_fasta(path, names, seqs, width=70):
    """70-column FASTA (the reference rejects any other wrap width, fasta.cpp:83-87)."""
    with open(path, "wb") as f:
        for name, s in zip(names, seqs):
            f.write(b">" + name.encode() + b"\n")
            b = np.asarray(s, dtype=np.uint8).tobytes()
            for k in range(0, len(b), width):
                f.write(b[k:k + width] + b"\n")
