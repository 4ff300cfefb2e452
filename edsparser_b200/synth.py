"""numpy statement of the synthetic-alignment generator (k_synth in csrc/msa.cu): BASELINE.json
configs 2 / 4. Used by bench.py's reference arm / cpu_baseline leg to write the SAME
alignment to a file without touching the GPU, and by the tests to pin the device generator."""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def alignment(n_rows, total_cols, col_begin=0, col_count=None, seed=1, variable_ppm=10000):
    """rows x col_count uint8 matrix of the window [col_begin, col_begin + col_count)."""
    if col_count is None:
        col_count = total_cols - col_begin
    with np.errstate(over="ignore"):
        cols = np.arange(col_begin, col_begin + col_count, dtype=np.uint64)
        hc = _mix64(np.uint64(seed) ^ _mix64(cols))
        base0 = _BASES[(hc & np.uint64(3)).astype(np.intp)]
        m = np.tile(base0, (n_rows, 1))
        var = np.nonzero(((hc >> np.uint64(8)) % np.uint64(1000000)) < np.uint64(variable_ppm))[0]
        hv = hc[var]
        for r in range(1, n_rows):
            hr = _mix64(hv ^ _mix64(np.uint64((r << 1) | 1)))
            roll = (hr >> np.uint64(8)) % np.uint64(100)
            sub = roll < np.uint64(30)
            gap = (roll >= np.uint64(30)) & (roll < np.uint64(40))
            row = m[r]
            row[var[sub]] = _BASES[(hr[sub] & np.uint64(3)).astype(np.intp)]
            row[var[gap]] = ord("-")
    return m


def fasta_window(n_rows, total_cols, line_width=80, col_begin=0, col_count=None, seed=1, variable_ppm=10000):
    """The exact bytes eds_msa_synth_device lays out for this window:
    '>seq<r+1>\\n' + the row's slice of the wrapped text + '\\n' per row."""
    if col_count is None:
        col_count = total_cols - col_begin
    m = alignment(n_rows, total_cols, col_begin, col_count, seed, variable_ppm)
    lw = line_width
    u_begin = col_begin + col_begin // lw
    last = col_begin + col_count - 1
    row_bytes = last + last // lw - u_begin + 1
    g = np.arange(col_begin, col_begin + col_count, dtype=np.int64)
    pos = g + g // lw - u_begin
    out = bytearray()
    seg = np.full(row_bytes, ord("\n"), dtype=np.uint8)
    for r in range(n_rows):
        out += b">seq%d\n" % (r + 1)
        seg[pos] = m[r]
        out += seg.tobytes()
        out += b"\n"
    return bytes(out)


def genrandomeds(ref_size, variability_ppm=100_000, paths=4, seed=1):
    """numpy statement of eds_genrandomeds_device (csrc/leds.cu GenFn): the same (seed, position)-keyed generator,
    position by position. Returns (eds bytes, seds bytes). For sizes a Python loop finishes (tests)."""
    M = (1 << 64) - 1

    def mix(x):
        x = (x + 0x9E3779B97F4A7C15) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)

    def site(i):
        h = mix(seed ^ mix(i))
        base = h & 3
        is_site = ((h >> 8) % 1000000) < variability_ppm
        return h, base, (min(paths, 2 + (h >> 40) % 3) if is_site else 0)

    def alt(h, base, k):
        hk = mix(h ^ mix(k))
        r = hk % 100
        if r < 70:
            return "ACGT"[(base + 1 + (hk >> 8) % 3) & 3]
        if r < 85:
            extra = 1 + (hk >> 16) % 10
            return "ACGT"[base] + "".join("ACGT"[mix((hk + j) & M) & 3] for j in range(extra))
        return ""

    eds, seds = [], []
    sites = [site(i) for i in range(ref_size)]
    for i, (h, base, n_alts) in enumerate(sites):
        if not n_alts:
            first = i == 0 or sites[i - 1][2] != 0
            last = i + 1 == ref_size or sites[i + 1][2] != 0
            if first:
                eds.append("{")
                seds.append("{0}")
            eds.append("ACGT"[base])
            if last:
                eds.append("}")
            continue
        alts = ["ACGT"[base]] + [alt(h, base, k) for k in range(1, n_alts)]
        eds.append("{" + ",".join(alts) + "}")
        owner = [p - 1 if p <= n_alts else mix(h ^ mix(100 + p)) % n_alts for p in range(1, paths + 1)]
        for k in range(n_alts):
            seds.append("{" + ",".join(str(p) for p in range(1, paths + 1) if owner[p - 1] == k) + "}")
    return "".join(eds).encode(), "".join(seds).encode()
