"""Seeded input generators for the parity tests (numpy only)."""
import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_alignment(rng, n_rows, n_cols, p_var=0.1, p_sub=0.3, p_gap=0.15, p_ref_gap=0.05, cluster=0.5,
                     alphabet=ALPHABET):
    """rows x cols uint8 matrix. Variable columns cluster (so runs of every length appear), row 0 may
    hold '-', rows may differ from row 0 by substitution or gap."""
    ref = alphabet[rng.integers(0, len(alphabet), n_cols)]
    m = np.tile(ref, (n_rows, 1))
    var = np.zeros(n_cols, dtype=bool)
    state = False
    for c in range(n_cols):
        state = rng.random() < (cluster if state else p_var)
        var[c] = state
    cols = np.nonzero(var)[0]
    for c in cols:
        if rng.random() < p_ref_gap:
            m[:, c] = ord("-")  # row 0 is a gap: column is not conserved even if every row agrees
        roll = rng.random(n_rows)
        sub = roll < p_sub
        gap = (roll >= p_sub) & (roll < p_sub + p_gap)
        sub[0] = gap[0] = False
        m[sub, c] = alphabet[rng.integers(0, len(alphabet), int(sub.sum()))]
        m[gap, c] = ord("-")
    return m


def to_fasta(m, wrap, headers=None, final_newline=True, blank_before_header=False):
    """Gapped FASTA text of the matrix, every row wrapped at `wrap` residues."""
    out = bytearray()
    for r in range(m.shape[0]):
        if blank_before_header and r > 0:
            out += b"\n"
        out += (headers[r] if headers else b">s%d" % (r + 1)) + b"\n"
        row = m[r].tobytes()
        for i in range(0, len(row), wrap):
            out += row[i:i + wrap] + b"\n"
    if not final_newline and out.endswith(b"\n"):
        out = out[:-1]
    return bytes(out)


def random_msa_text(rng, max_rows=9, max_cols=200, **kw):
    n_rows = int(rng.integers(2, max_rows + 1))
    n_cols = int(rng.integers(1, max_cols + 1))
    wrap = int(rng.choice([1, 2, 3, 7, 16, 31, 32, 33, 60, 80, n_cols, n_cols + 5]))
    wrap = max(1, wrap)
    m = random_alignment(rng, n_rows, n_cols, p_var=float(rng.choice([0.0, 0.02, 0.1, 0.4])),
                         cluster=float(rng.choice([0.0, 0.5, 0.8])), **kw)
    headers = [b">" + bytes(rng.integers(97, 123, int(rng.integers(0, 20))).astype(np.uint8)) for _ in range(n_rows)]
    text = to_fasta(m, wrap, headers, final_newline=bool(rng.integers(0, 2)))
    return text, m, wrap


def random_eds(rng, n_sym=None, paths=4, p_deg=0.4, max_common=14, with_sources=True, universal_rate=0.1, compact_in=False):
    """Random EDS text (+ SEDS text): degenerate symbols with 2-4 alternatives (some empty), commons of
    random length. Sources: every alternative gets a non-empty subset of 1..paths (subsets of one symbol
    cover all paths, may overlap), commons get {0}; a few sets carry the universal marker 0 next to ids."""
    if n_sym is None:
        n_sym = int(rng.integers(1, 30))
    eds, seds = [], []
    for _ in range(n_sym):
        if rng.random() < p_deg:
            k = int(rng.integers(2, 5))
            alts = []
            for _a in range(k):
                ln = int(rng.choice([0, 1, 1, 2, 3, 6]))
                alts.append(bytes(ALPHABET[rng.integers(0, 4, ln)]))
            eds.append(b"{" + b",".join(alts) + b"}")
            owner = rng.integers(0, k, paths)  # every path picks one alternative
            for a in range(k):
                ids = set(int(p) + 1 for p in np.nonzero(owner == a)[0])
                if rng.random() < 0.3:
                    ids.add(int(rng.integers(1, paths + 1)))  # overlap (heterozygous)
                if not ids:
                    ids.add(int(rng.integers(1, paths + 1)))
                if rng.random() < universal_rate:
                    ids.add(0)
                seds.append(b"{" + b",".join(b"%d" % i for i in sorted(ids, key=lambda _x: rng.random())) + b"}")
        else:
            ln = int(rng.integers(0 if rng.random() < 0.05 else 1, max_common + 1))
            s = bytes(ALPHABET[rng.integers(0, 4, ln)])
            eds.append(s if (compact_in and ln > 0) else b"{" + s + b"}")
            seds.append(b"{0}" if rng.random() > 0.1 else b"{%d}" % int(rng.integers(1, paths + 1)))
    return b"".join(eds), (b"".join(seds) if with_sources else None)
