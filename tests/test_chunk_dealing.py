"""The dealt / stolen chunk numbering of a shared frame (next_chunk, cobbletrace_b200/csrc/ct_kernels.cuh), restated in Python and checked
exhaustively on the CPU: for every participant count R, dealt fraction E/8, run length 2^rs and chunk count, the participants' dealt
entries plus the entries stolen from the one cursor must name every chunk of the tile exactly once.  (The device code itself is
exercised by tests/test_gpu_parity.py::test_shared_frame_partition_covers_the_frame_once and tests/test_multi_gpu_ipc.py; this
restatement covers the parameter space those cannot afford.)"""
import itertools


def dealt_chunks(n_chunks, R, E, rs, part):
    """Participant `part`'s dealt share: the loop over static_next in next_chunk."""
    L = 1 << rs
    n_runs = (n_chunks + L - 1) >> rs
    G = 8 * R
    n_groups = (n_runs + G - 1) // G
    out = []
    c = 0
    while E > 0:
        cr, off = c >> rs, c & (L - 1)
        g = cr // E
        if g >= n_groups:
            break
        idx = ((g * G + (cr - g * E) * R + part) << rs) + off
        if idx < n_chunks:
            out.append(idx)
        c += 1
    return out


def stolen_chunks(n_chunks, R, E, rs):
    """Everything the shared cursor hands out (to whoever asks)."""
    L = 1 << rs
    n_runs = (n_chunks + L - 1) >> rs
    G = 8 * R
    n_groups = (n_runs + G - 1) // G
    per = (8 - E) * R
    out = []
    d = 0
    while per:
        dr, off = d >> rs, d & (L - 1)
        g = dr // per
        if g >= n_groups:
            break
        idx = ((g * G + E * R + (dr - g * per)) << rs) + off
        if idx < n_chunks:
            out.append(idx)
        d += 1
    return out


def emulated_share(n_chunks, stride, rs):
    """Option "emulate_ranks": the runs rank 0 of `stride` GPUs would own."""
    L = 1 << rs
    out = []
    c = 0
    while True:
        run = (c >> rs) * stride
        if (run << rs) >= n_chunks:
            return out
        idx = (run << rs) + (c & (L - 1))
        if idx < n_chunks:
            out.append(idx)
        c += 1


def test_every_chunk_is_dealt_or_stolen_exactly_once():
    sizes = [1, 2, 7, 31, 32, 33, 63, 64, 65, 127, 500, 1023, 1024, 1025, 4099]
    for R, E, rs in itertools.product((2, 3, 4, 5, 8, 16), range(0, 9), (0, 1, 2, 3, 5, 8)):
        for n in sizes:
            got = stolen_chunks(n, R, E, rs)
            for part in range(R):
                got += dealt_chunks(n, R, E, rs, part)
            assert sorted(got) == list(range(n)), (R, E, rs, n)


def test_dealt_shares_are_balanced_and_interleaved():
    n, R, E = 8 * 8 * 40, 8, 7
    for rs in (0, 2, 4):
        shares = [dealt_chunks(n, R, E, rs, p) for p in range(R)]
        assert max(map(len, shares)) - min(map(len, shares)) <= (1 << rs)
        for p, sh in enumerate(shares):                       # runs of 2^rs consecutive chunks, R runs apart
            runs = sorted({c >> rs for c in sh})
            assert all((r % (8 * R)) % R == p for r in runs if (r % (8 * R)) < E * R), (rs, p)


def test_emulated_share_is_rank_zeros_runs():
    for stride, rs, n in itertools.product((2, 4, 8), (0, 1, 3, 6), (1, 31, 64, 1000, 4099)):
        got = emulated_share(n, stride, rs)
        want = [c for c in range(n) if ((c >> rs) % stride) == 0]
        assert got == want, (stride, rs, n)
