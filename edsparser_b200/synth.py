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
