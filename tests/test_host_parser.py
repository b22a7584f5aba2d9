"""The host reader of the drop-in drivers (host/gt_host.cpp: producer-thread line reader + multi-threaded BED/REG/GFF
parsers) against the reference's own reader, on a CPU-only box: `bin/gt_regdump FILE` prints what our parser makes
of a file in REG spelling, `oracle/_ref/genomic_regions reg FILE` does the same through the reference's
GenomicRegionSet (genomic_intervals.cpp:3667-3914).  Run with 1 and several parsing threads and with pieces of a
few hundred bytes, so that every line boundary is a piece boundary somewhere."""
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

import support

BIN = os.path.join(support.ROOT, "ibm-cbc-genomic-tools_b200", "bin")
DUMP = os.path.join(BIN, "gt_regdump")

pytestmark = pytest.mark.skipif(not os.path.exists(DUMP), reason="bin/gt_regdump not built")

THREADINGS = [{"GT_PARSE_THREADS": "1"}, {"GT_PARSE_THREADS": "5", "GT_PARSE_PIECE_BYTES": "300"}]


def dump(path, env_extra, args=(), stdin=None):
    env = dict(os.environ, **env_extra)
    p = subprocess.run([DUMP] + list(args) + [str(path)], input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    return p.returncode, p.stdout, p.stderr


def strip_extras(out):
    """gt_regdump appends `\\tw=..` (with -w) and `\\t#line`; the reference prints label and intervals only"""
    return b"".join(re.sub(rb"(\tw=-?\d+)?\t#\d+$", b"", l) + b"\n" for l in out.split(b"\n") if l)


def ref_reg(path, stdin=None):
    if not support.have_ref():
        pytest.skip("reference binaries not built (oracle/_ref)")
    return support.run_ref("genomic_regions", ["reg", str(path)], stdin=stdin, check=False)


def rand_lines(rng, n, kind):
    names = ["chr1", "chr10", "chr2", "chrX", "chrUn_gl000220", "scaffold_12345678"]
    out = []
    for k in range(n):
        c = names[rng.integers(len(names))]
        s = int(rng.integers(0, 5_000_000))
        e = s + int(rng.integers(1, 3000))
        strand = "+-"[rng.integers(2)]
        lab = "L%d" % k if rng.integers(4) else str(int(rng.integers(-3, 40)))
        if kind == "bed3":
            out.append(f"{c}\t{s}\t{e}")
        elif kind == "bed4":
            out.append(f"{c}\t{s}\t{e}\t{lab}")
        elif kind == "bed5":
            out.append(f"{c}\t{s}\t{e}\t{lab}\t0")
        elif kind == "bed6":
            st = [strand, strand, ".", "1", "-1"][rng.integers(5)]
            out.append(f"{c}\t{s}\t{e}\t{lab}\t{rng.integers(1000)}\t{st}")
        elif kind == "bed6_space":
            out.append(f"{c} {s} {e} {lab} 0 {strand}")
        elif kind == "bed_mixed":                                       # column count varies line by line; BED12 blocks among them
            if rng.integers(3) == 0:
                nb = int(rng.integers(1, 5))
                sizes = [int(rng.integers(1, 50)) for _ in range(nb)]
                starts = [i * 100 for i in range(nb)]
                out.append("\t".join([c, str(s), str(s + starts[-1] + sizes[-1]), lab, "0", strand, str(s), str(e), "0", str(nb),
                                      ",".join(map(str, sizes)), ",".join(map(str, starts))]))
            else:
                cols = [c, str(s), str(e), lab, "0", strand, str(s), str(e), "0"][:int(rng.integers(3, 10))]
                if len(cols) == 7:
                    cols = cols[:6]
                out.append("\t".join(cols))
        elif kind == "bed_odd":                                         # things the one-pass path must hand to the general one
            pick = rng.integers(6)
            if pick == 0:
                out.append(f"{c}\t {s}\t{e}\t{lab}\t0\t{strand}")       # blank before a number
            elif pick == 1:
                out.append(f"{c}\t{s}x\t{e}abc\t{lab}\t0\t{strand}")    # atol stops at the first non-digit
            elif pick == 2:
                out.append(f"{c}\t{s}\t{e}\t my label\t0\t{strand}")    # leading blank in the label is skipped
            elif pick == 3:
                out.append(f"{c}\t-{s}\t{e}\t{lab}\t0\t{strand}")       # negative start
            elif pick == 4:
                out.append(f"{c}\t{s}\t{e}\t{lab}\t0\t{strand}\r")      # CR before the newline: 6th column is '+\r'?  (fatal or not: same as reference)
            else:
                out.append(f"{c}\t{s}\t{e}\t{lab}\t\t{strand}")         # empty score column
        elif kind == "reg":
            nb = int(rng.integers(1, 4))
            iv = " ".join(f"{c} {strand} {s + 500 * i + 1} {s + 500 * i + 100}" for i in range(nb))
            out.append(f"{lab}\t{iv}")
        elif kind == "reg_compact":
            nb = int(rng.integers(1, 4))
            out.append(f"{lab}\t{c} {strand} " + ",".join(str(s + 500 * i + 1) for i in range(nb)) + " " + ",".join(str(s + 500 * i + 100) for i in range(nb)))
        elif kind == "sam":
            # POS + CIGAR: M/D/X consume the reference, N splits into blocks, I/S add to the fragment only, H/P to neither
            ops, frag = [], 0
            for k_op in range(int(rng.integers(1, 6))):
                op = "M" if k_op == 0 else "MMMMIDNSHXP"[rng.integers(11)]   # (a read without a reference-consuming operation has no interval at all)
                ln = int(rng.integers(1, 60))
                ops.append("%d%s" % (ln, op))
                if op in "MISX":
                    frag += ln
            cigar = "".join(ops)
            pick = rng.integers(5)
            seq = "*" if pick == 0 else "A" * frag if frag else "*"
            if pick == 1:
                cigar, seq = "*", "ACGT" * int(rng.integers(1, 20))       # "*" stands for <length of SEQ>M
            flag = int(rng.integers(0, 4096))
            cols = [lab, str(flag), c, str(s + 1), "60", cigar, "*", "0", "0", seq, "*"] + (["NM:i:0", "XS:A:+"] if rng.integers(2) else [])
            out.append("\t".join(cols))
        elif kind == "gff":
            st = [strand, "."][rng.integers(2)]
            cols = [c, "src", "feat", str(s + 1), str(e), ".", st, ".", lab][:int(rng.integers(8, 10))]
            out.append("\t".join(cols))
    return out


KINDS = ["bed3", "bed4", "bed5", "bed6", "bed6_space", "bed_mixed", "reg", "reg_compact", "gff", "sam"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("env", THREADINGS, ids=["t1", "t5"])
def test_parser_matches_reference_reader(tmp_path, kind, env):
    rng = np.random.default_rng(abs(hash(kind)) % 1000 + 7)
    lines = rand_lines(rng, 3000, kind)
    path = tmp_path / ("in." + kind)
    path.write_text("\n".join(lines) + "\n")
    want = ref_reg(path)
    got = dump(path, env)
    assert want[0] == 0 and got[0] == 0, (want[2][-200:], got[2][-200:])
    assert strip_extras(got[1]) == want[1]
    # line numbers: region k is on line k + 1
    nums = [int(l.rsplit(b"#", 1)[1]) for l in got[1].splitlines()]
    assert nums == list(range(1, len(lines) + 1))


@pytest.mark.parametrize("env", THREADINGS, ids=["t1", "t5"])
def test_odd_bed_lines_one_by_one(tmp_path, env):
    """each odd line alone after a few clean ones: same regions, or the same fatal message and exit code"""
    rng = np.random.default_rng(5)
    clean = rand_lines(rng, 40, "bed6")
    for k, odd in enumerate(rand_lines(rng, 60, "bed_odd")):
        path = tmp_path / f"odd{k}.bed"
        path.write_text("\n".join(clean[:20] + [odd] + clean[20:]) + "\n")
        want = ref_reg(path)
        got = dump(path, env)
        assert got[0] == want[0], (odd, got[2], want[2])
        if want[0] == 0:
            assert strip_extras(got[1]) == want[1], odd
        else:
            assert got[2] == want[2], odd


@pytest.mark.parametrize("env", THREADINGS, ids=["t1", "t5"])
def test_headers_gzip_stdin_and_unterminated_last_line(tmp_path, env):
    rng = np.random.default_rng(11)
    body = rand_lines(rng, 500, "bed6")
    text = "track name=x\nbrowser position chr1\n" + "\n".join(body) + "\nchr1\t5\t9\tdropped\t0\t+"      # no final newline
    path = tmp_path / "h.bed"
    path.write_text(text)
    want = ref_reg(path)
    got = dump(path, env)
    assert got[0] == 0 and strip_extras(got[1]) == want[1]
    assert b"dropped" not in got[1]
    first = got[1].splitlines()[0]
    assert first.endswith(b"#3")                                         # two header lines come first
    with gzip.open(str(path) + ".gz", "wb") as g:
        g.write(text.encode())
    got_gz = dump(str(path) + ".gz", env)
    assert got_gz[1] == got[1]
    got_in = dump("-", env, stdin=text.encode())
    assert got_in[1] == got[1]
    sam = "@HD\tVN:1.0\n@SQ\tSN:chr1\tLN:1000\n" + "\n".join(rand_lines(rng, 80, "sam")) + "\n"
    p3 = tmp_path / "h.sam"
    p3.write_text(sam)
    assert strip_extras(dump(p3, env)[1]) == ref_reg(p3)[1]
    gff = "##gff-version 3\n##x\n" + "\n".join(rand_lines(rng, 50, "gff")) + "\n"
    p2 = tmp_path / "h.gff"
    p2.write_text(gff)
    assert strip_extras(dump(p2, env)[1]) == ref_reg(p2)[1]
    empty = tmp_path / "empty.bed"
    empty.write_text("")
    assert dump(empty, env)[:2] == (0, b"")


@pytest.mark.parametrize("env", THREADINGS, ids=["t1", "t5"])
@pytest.mark.parametrize("bad,kind", [("chr1\t5", "bed6"), ("chr1\t5\t9\tx\t0\t*", "bed6"), ("lab\tchr1 + 5", "reg"),
                                      ("lab\tchr1 + 1,2 5", "reg_compact"), ("chr1\ta\tb\t1\t2\t.\t+", "gff"),
                                      ("r\t0\tchr1\t5\t60\t10M\t*\t0\t0\tACGT", "sam"),                       # 10 columns
                                      ("r\t0\tchr1\t5\t60\t4=\t*\t0\t0\tACGT\t*", "sam"),                    # '=' is not in the tokenizer's alphabet
                                      ("r\t0\tchr1\t5\t60\t3M2Z\t*\t0\t0\tACGTA\t*", "sam"),
                                      ("r7\t16\tchr1\t5\t60\t3M2I\t*\t0\t0\tACGTACGT\t*", "sam")])           # SEQ length != CIGAR's
def test_first_malformed_line_wins(tmp_path, env, bad, kind):
    """several malformed lines, parsed by different threads: the earliest is reported, with the reference's words;
    the regions before it are still delivered (the drivers feed them to the engine before failing)"""
    rng = np.random.default_rng(3)
    lines = rand_lines(rng, 900, kind)
    lines[400] = bad
    lines[650] = bad
    lines[880] = bad
    path = tmp_path / "bad.txt"
    path.write_text("\n".join(lines) + "\n")
    want = ref_reg(path)
    got = dump(path, env)
    assert want[0] != 0 and got[0] == want[0]
    assert got[2] == want[2]
    assert len(got[1].splitlines()) == 400


@pytest.mark.parametrize("env", THREADINGS, ids=["t1", "t5"])
def test_label_weights_and_chunked_reads(tmp_path, env):
    rng = np.random.default_rng(17)
    lines = rand_lines(rng, 2000, "bed6")
    path = tmp_path / "w.bed"
    path.write_text("\n".join(lines) + "\n")
    whole = dump(path, env, args=["-w", "25"])
    chunked = dump(path, dict(env, GT_PARSE_PIECE_BYTES="200"), args=["-w", "25", "-c", "100"])
    assert whole[0] == 0 and whole[1] == chunked[1]
    for l, src in zip(whole[1].splitlines(), lines):
        lab = src.split("\t")[3]
        try:
            v = int(lab)
        except ValueError:
            v = 0                                                        # atol of a non-number
        assert l.split(b"\t")[2] == b"w=%d" % min(25, v)                 # GetLabelValue, genomic_intervals.cpp:1081-1085


def test_thread_counts_agree_on_a_large_file(tmp_path):
    synth = os.path.join(BIN, "gt_synth_bed")
    if not os.path.exists(synth):
        pytest.skip("bin/gt_synth_bed not built")
    path = tmp_path / "reads.bed"
    with open(path, "wb") as f:
        subprocess.run([synth, "300000", "2"], stdout=f, check=True)
    a = dump(path, {"GT_PARSE_THREADS": "1"})
    b = dump(path, {"GT_PARSE_THREADS": "7", "GT_PARSE_PIECE_BYTES": "65536"})
    assert a[0] == 0 and a[1] == b[1] and len(a[1].splitlines()) == 300000
    # and the stream is the one the device generator and tests/support.py produce
    r = support.synth_reads(1000, 2)
    c, s, e, st = r["chrom"], r["start"], r["stop"], r["strand"]
    for k, l in enumerate(a[1].splitlines()[:1000]):
        f = l.split(b"\t")[1].split(b" ")
        assert f[0].decode() == support.HG19_NAMES[c[k]] and f[1] == bytes([int(st[k])]) and int(f[2]) == s[k] and int(f[3]) == e[k]


def test_blocks_and_long_lines(tmp_path):
    """A file that spans several of the reader's 32 MB blocks (lines cut by block ends are carried over), and a line longer
    than the reader's first block (the block grows)."""
    synth = os.path.join(BIN, "gt_synth_bed")
    if not os.path.exists(synth):
        pytest.skip("bin/gt_synth_bed not built")
    path = tmp_path / "big.bed"
    n = 2_000_000                                                        # ~72 MB of text
    with open(path, "wb") as f:
        subprocess.run([synth, str(n), "9"], stdout=f, check=True)
    assert os.path.getsize(path) > 2 * (32 << 20)
    for env in ({"GT_PARSE_THREADS": "1"}, {"GT_PARSE_THREADS": "6"}):
        rc, out, err = dump(path, env, args=["-q"])
        assert rc == 0 and out.split()[-4:] == [b"regions", b"%d" % n, b"intervals", b"%d" % n], (out, err)
    # last lines of the file through the full dump: line numbers still right after the block boundaries
    tail = subprocess.run(["tail", "-n", "3", str(path)], stdout=subprocess.PIPE, check=True).stdout
    small = tmp_path / "tail.bed"
    small.write_bytes(tail)
    want = strip_extras(dump(small, {"GT_PARSE_THREADS": "1"})[1])
    got = dump(path, {"GT_PARSE_THREADS": "6"}, args=["-c", "500000"])[1].splitlines()[-3:]
    assert strip_extras(b"\n".join(got) + b"\n") == want
    assert [int(l.rsplit(b"#", 1)[1]) for l in got] == [n - 2, n - 1, n]
    # a 3 MB label in the middle of a small file
    rng = np.random.default_rng(23)
    lines = rand_lines(rng, 50, "bed6")
    lines[20] = "chr1\t100\t200\t" + "x" * (3 << 20) + "\t0\t-"
    longp = tmp_path / "long.bed"
    longp.write_text("\n".join(lines) + "\n")
    want = ref_reg(longp)
    for env in THREADINGS:
        got = dump(longp, env)
        assert got[0] == 0 == want[0] and strip_extras(got[1]) == want[1]


@pytest.mark.parametrize("kind,n_cases", [("bed6", 250), ("gff", 100), ("reg", 100), ("reg_compact", 60), ("sam", 120)])
def test_mutated_lines_differential(tmp_path, kind, n_cases):
    """Differential fuzz against the reference's reader (for BED: of the one-pass path and its hand-over to the general one):
    clean lines with one to three random edits (inserted / replaced / deleted characters from an alphabet of separators,
    signs, digits, CIGAR letters, CR).  Each mutated line sits alone between clean lines; regions or fatal message + exit
    code must be the reference's."""
    rng = np.random.default_rng(2024)
    clean = rand_lines(rng, 30, kind)
    alphabet = list("\t \r+-.1a0,;MN*=")
    mismatches = []
    for k in range(n_cases):
        line = list(rand_lines(rng, 1, kind)[0])
        for _ in range(int(rng.integers(1, 4))):
            pos = int(rng.integers(0, len(line) + 1))
            what = rng.integers(3)
            ch = alphabet[rng.integers(len(alphabet))]
            if what == 0 or not line:
                line.insert(pos, ch)
            elif what == 1:
                line[min(pos, len(line) - 1)] = ch
            else:
                del line[min(pos, len(line) - 1)]
        mutated = "".join(line)
        if "\n" in mutated or not mutated.strip():
            continue
        path = tmp_path / ("m%d.bed" % k)
        path.write_text("\n".join(clean[:10] + [mutated] + clean[10:]) + "\n")
        want = ref_reg(path)
        got = dump(path, {"GT_PARSE_THREADS": "2", "GT_PARSE_PIECE_BYTES": "200"})
        ok = got[0] == want[0] and (strip_extras(got[1]) == want[1] if want[0] == 0 else got[2] == want[2])
        if b"does not fit in 32 bits" in got[2]:
            continue                                                     # the documented divergence: the SoA of the C ABI is 32-bit
        if b"unknown CIGAR operation type" in got[2] and b"unknown CIGAR operation type" in want[2] and got[0] == want[0]:
            continue                                                     # a CIGAR ending in digits: the reference reads past the string's end for the "type"
        if not ok:
            mismatches.append((mutated, want[0], got[0], want[2][-120:], got[2][-120:]))
    assert not mismatches, mismatches[:5]


# ------------------------------------------------------------------------------------------------
# gt::PrintRegion (what subset / overlap / gsort print) against GenomicRegion*::Print of the reference, without a GPU:
# `genomic_overlaps subset -inv` against a reference set on a chromosome no query is on prints every query through Print()
# ------------------------------------------------------------------------------------------------
def test_print_region_matches_reference_print(tmp_path):
    if not support.have_ref():
        pytest.skip("reference binaries not built (oracle/_ref)")
    (tmp_path / "far.bed").write_text("chrFAR\t1\t2\tx\t0\t+\n")
    rng = np.random.default_rng(5)
    cases = {
        "odd.bed": "chr1 100 200 a 5 +\nchr1 100 200\nchr1 100 200 b\nchr1 100 200 c 7\nchr1\t5\t50\td\t3.7\t-\t7\t9\n"
                   "chr1\t5\t50\td\t12\t.\t7\t9\t255,0,0\nchr1\t5\t50\td\t12\t1\t7\t9\t255,0,0\t2\n"
                   "chr1\t1000\t2000\tgD\t0\t+\t1000\t2000\t0\t2\t100,100\t0,900\n",
        "v.reg": "r1\tchr1 + 100 200 chr1 + 500 600\nr2\tchr1 - 150 160\nr3\tchr2 + 100,300 150,400\n",
        "v.gff": "##gff-version 2\nchr1\tsrc\tgene\t11\t20\t.\t-\t.\tgC\nchr1\tsrc\tgene\t5\t20\t0.5\t+\t2\tgD\tnote here\nchr1\tsrc\tgene\t5\t20\t.\t.\t.\n",
        "track.bed": "track name=x\nbrowser position chr1\nchr1\t5\t9\tz\t1\t-\n",
    }
    sam = ["@HD\tVN:1.0", "@SQ\tSN:chr1\tLN:5000"]
    for k in range(500):
        ln = int(rng.integers(1, 90))
        cig = ("%dM" % ln) if k % 7 else "*"
        seq = "A" * (ln if k % 7 else 36)
        sam.append("\t".join(["r%d" % k, str(int(rng.integers(0, 2048))), "chr1", str(int(rng.integers(1, 4000))), "60", cig, "=", "7", "-3", seq, "*"] +
                             (["NM:i:1", "XS:A:+"] if k % 3 == 0 else [])))
    cases["q.sam"] = "\n".join(sam) + "\n"
    for name, text in cases.items():
        (tmp_path / name).write_text(text)
        rc, want, err = support.run_ref("genomic_overlaps", ["subset", "-inv", tmp_path / "far.bed", tmp_path / name], check=False)
        assert rc == 0 and len(want) > 0, (name, err)
        rc2, got, err2 = dump(tmp_path / name, {}, args=("-p",))
        assert rc2 == 0 and got == want, (name, got[:300], want[:300])


# ------------------------------------------------------------------------------------------------
# BAM input (FileBufferBAM, core.cpp:371-430): records re-spelt as SAM lines, then the SAM reader
# ------------------------------------------------------------------------------------------------
def _bam_case(tmp_path, with_header=True, n=400, seed=11):
    import struct
    rng = np.random.default_rng(seed)
    refs = [("chr1", 50_000), ("chr2", 40_000), ("chrUn_gl000220", 9_000)]
    recs = []
    for k in range(n):
        ln = int(rng.integers(1, 80))
        spliced = k % 5 == 0 and ln > 10
        cigar = [(ln // 2, "M"), (int(rng.integers(50, 900)), "N"), (ln - ln // 2, "M")] if spliced else ([(2, "S"), (ln, "M")] if k % 11 == 3 else [(ln, "M")])
        seq_len = sum(c for c, op in cigar if op in "MIS=X")
        aux = b""
        if k % 3 == 0:
            aux += b"NMC" + bytes([k % 7]) + b"XSA+" + b"MDZ" + b"%dA3" % (k % 9) + b"\x00"
        if k % 10 == 1:
            aux += b"ASi" + struct.pack("<i", -k) + b"XFf" + struct.pack("<f", 0.25 * k) + b"ZBBs" + struct.pack("<ihh", 2, -3, 7) + b"UQS" + struct.pack("<H", 40000)
        unmapped = k % 37 == 36
        recs.append({"qname": "read%d" % k, "flag": int(rng.integers(0, 2048)) & ~4 | (4 if unmapped else 0), "tid": -1 if unmapped else int(rng.integers(0, 3)),
                     "pos": -1 if unmapped else int(rng.integers(0, 8000)), "mapq": int(rng.integers(0, 61)),
                     "cigar": [] if (unmapped or k % 13 == 5) else cigar, "mtid": [-1, 0, 1][k % 3], "mpos": int(rng.integers(0, 8000)), "isize": int(rng.integers(-500, 500)),
                     "seq": "" if k % 17 == 9 else "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, seq_len)),
                     "qual": None if k % 4 == 0 else [int(x) for x in rng.integers(0, 42, seq_len)], "aux": aux})
        if recs[-1]["seq"] == "":
            recs[-1]["qual"] = None
    header = "@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:chr1\tLN:50000\n@SQ\tSN:chr2\tLN:40000\n@SQ\tSN:chrUn_gl000220\tLN:9000\n@PG\tID:x" if with_header else ""
    path = tmp_path / ("h.bam" if with_header else "n.bam")
    support.write_bam(path, header, refs, recs)
    return path


@pytest.mark.parametrize("with_header", [True, False])
def test_bam_input_matches_reference_reader(tmp_path, with_header):
    """regions of a BAM file (spliced reads, soft clips, '*' CIGARs, unmapped reads, every aux type) and every line as Print()
    writes it back, against the reference built with its vendored samtools"""
    if not support.have_ref():
        pytest.skip("reference binaries not built (oracle/_ref)")
    path = _bam_case(tmp_path, with_header)
    rc, want, err = ref_reg(path)
    for env in THREADINGS:
        rc2, got, err2 = dump(path, env)
        assert (rc2, strip_extras(got)) == (rc, want), (got[:400], want[:400], err, err2)
        assert rc != 0 or len(want) > 100
    (tmp_path / "far.bed").write_text("chrFAR\t1\t2\tx\t0\t+\n")
    rc, want, err = support.run_ref("genomic_overlaps", ["subset", "-inv", tmp_path / "far.bed", path], check=False)
    rc2, got, err2 = dump(path, {}, args=("-p",))
    assert (rc2, got) == (rc, want), (got[:400], want[:400], err[-300:], err2[-300:])


def test_bed_fast_path_word_boundaries(tmp_path):
    """The clean-BED path reads whole 8-byte words (chromosome key, eight digits at a time, field ends): chromosome names of every
    length around 8, numbers of 1 to 18 digits with and without leading zeros, labels and scores of every length around the word
    size, every strand spelling, 3 to 6 columns -- against the reference's reader, in one piece and in many small ones."""
    rng = np.random.default_rng(808)
    chroms = ["c", "c2", "chr", "chr1", "chr10", "chrUn_g", "chrUn_gl", "chrUn_gl0", "12345678", "123456789", "scaffold_123456789"]
    lines = []
    for k in range(40000):
        c = chroms[rng.integers(len(chroms))]
        nd = int(rng.integers(1, 10))
        a = int(rng.integers(10 ** (nd - 1) if nd > 1 else 0, 10 ** nd))
        b = a + int(rng.integers(0, 2000))
        sa, sb = str(a), str(b)
        if rng.random() < 0.1:
            sa = "0" * int(rng.integers(1, 19 - len(sa))) + sa
        if rng.random() < 0.1:
            sb = "0" * int(rng.integers(1, 20 - len(sb))) + sb                 # up to 19 characters: one more than the fast path takes
        t = [c, sa, sb]
        cols = int(rng.integers(3, 7))
        if cols >= 4:
            t.append("L" * int(rng.integers(1, 20)) if rng.random() < 0.5 else str(int(rng.integers(-5, 100000))))
        if cols >= 5:
            t.append("9" * int(rng.integers(1, 18)) if rng.random() < 0.5 else "0.%d" % k)
        if cols >= 6:
            t.append(["+", "-", ".", "1", "-1"][rng.integers(5)])
        lines.append("\t".join(t))
    lines += ["chr1\t1999999999\t2000000100\ta\t0\t+", "chr1\t2147483646\t2147483647", "chr1\t99999999\t100000000\tx", "chr1\t12345678\t123456789\tx\t5\t-1"]
    path = tmp_path / "words.bed"
    path.write_text("\n".join(lines) + "\n")
    want = ref_reg(path)
    assert want[0] == 0
    for env in ({"GT_PARSE_THREADS": "1"}, {"GT_PARSE_THREADS": "5", "GT_PARSE_PIECE_BYTES": "777"}):
        got = dump(path, env)
        assert got[0] == 0 and strip_extras(got[1]) == want[1], env


def bgzip(data, path, block=65280, level=6, eof_marker=True):
    """blocked gzip as bgzip writes it (SAM specification 4.1): members of at most 64 KB, each with its size in a 'BC' field"""
    import struct
    import zlib

    def member(d):
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = c.compress(d) + c.flush()
        return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                struct.pack("<II", zlib.crc32(d) & 0xFFFFFFFF, len(d)))
    with open(path, "wb") as f:
        for lo in range(0, len(data), block):
            f.write(member(data[lo:lo + block]))
        if eof_marker:
            f.write(member(b""))


def test_bgzf_members_inflated_side_by_side(tmp_path):
    """bgzip'd text and BAM are inflated member by member on several threads (LineReader::Bgzf), BAM records spelt as SAM lines on
    several threads: the same regions as the plain file gives, as zlib's gzread gives on one thread (GT_NO_BGZF=1) and as the
    reference's reader gives -- for whole files, for files cut inside a member, inside a header and at a member's end, and for
    stretches of more members than one round takes."""
    rng = np.random.default_rng(606)
    lines = "".join("chr%d\t%d\t%d\tr%d\t%d\t%s\n" % (rng.integers(1, 23), s, s + 50, k, k % 1000, "+-"[k % 2])
                    for k, s in enumerate(rng.integers(0, 200_000_000, 700_000))).encode()
    plain = tmp_path / "reads.bed"
    plain.write_bytes(lines)
    bgzip(lines, tmp_path / "reads.bed.gz", block=3000, level=1)                      # ~9 000 members, 17 MB of members: two rounds
    want = dump(plain, {})
    assert want[0] == 0
    for env in ({}, {"GT_INFLATE_THREADS": "1"}, {"GT_INFLATE_THREADS": "3", "GT_PARSE_THREADS": "2"}, {"GT_NO_BGZF": "1"}):
        got = dump(tmp_path / "reads.bed.gz", env)
        assert got[:2] == want[:2], env
    whole = (tmp_path / "reads.bed.gz").read_bytes()
    for cut in (len(whole) // 3, len(whole) // 3 + 5, 1_000_003, 18, 40, len(whole) - 28, len(whole) - 27):
        (tmp_path / "cut.bed.gz").write_bytes(whole[:cut])
        a, b = dump(tmp_path / "cut.bed.gz", {}), dump(tmp_path / "cut.bed.gz", {"GT_NO_BGZF": "1"})
        assert a[:2] == b[:2], cut
        if support.have_ref() and cut < 2_000_000:
            rc, ref_out, _ = ref_reg(tmp_path / "cut.bed.gz")
            assert (a[0], strip_extras(a[1])) == (rc, ref_out), cut
    # BAM: enough records for several formatting threads, and the file cut inside a member
    path = _bam_case(tmp_path, True, n=30_000, seed=12)
    a, b = dump(path, {}), dump(path, {"GT_NO_BGZF": "1", "GT_PARSE_THREADS": "1"})
    assert a[0] == 0 and a[:2] == b[:2] and a[1].count(b"\n") > 25_000
    if support.have_ref():
        rc, ref_out, _ = ref_reg(path)
        assert (a[0], strip_extras(a[1])) == (rc, ref_out)
    whole = path.read_bytes()
    for cut in (len(whole) // 2, len(whole) // 2 + 11):
        (tmp_path / "cut.bam").write_bytes(whole[:cut])
        a, b = dump(tmp_path / "cut.bam", {}), dump(tmp_path / "cut.bam", {"GT_NO_BGZF": "1"})
        assert a[:2] == b[:2] and a[1].count(b"\n") > 5_000, cut


def test_bam_record_longer_than_a_stretch(tmp_path):
    """a record of 9.5 MB (longer than the 8 MB of inflated stream the decoder cuts at a time) between ordinary ones"""
    refs = [("chr1", 100_000_000)]
    recs = [{"qname": "q%d" % k, "flag": 0, "tid": 0, "pos": 1000 + k, "mapq": 60, "cigar": [(50, "M")], "seq": "A" * 50, "qual": None} for k in range(3000)]
    big = 9_500_000
    recs.insert(1500, {"qname": "giant", "flag": 16, "tid": 0, "pos": 5000, "mapq": 1, "cigar": [(big, "M")], "seq": "C" * big, "qual": None})
    support.write_bam(tmp_path / "giant.bam", "@HD\tVN:1.0\n", refs, recs, block_bytes=60000)
    rc, got, _ = dump(tmp_path / "giant.bam", {})
    assert rc == 0 and got.count(b"\n") == 3001 and b"giant\tchr1 - 5001 9505000" in got
    if support.have_ref():
        rc, want, _ = ref_reg(tmp_path / "giant.bam")
        assert (0, strip_extras(got)) == (rc, want)


def test_sam_fast_path_shapes(tmp_path):
    """SAM lines of every clean shape the one-pass SAM path takes, and of shapes it must leave to the general tokeniser (blank-led
    fields, more than 16 blocks, names longer than 63 bytes): reference names of 1 to 70 bytes, positions of 1 to 9 digits and 0,
    CIGARs over M I D N S H P X with one to twenty operations or "*", SEQ "*" or of the CIGAR's length, 0 to 3 optional fields --
    against the reference's reader, in one piece and in many small ones."""
    rng = np.random.default_rng(515)
    names = ["c", "chr1", "chr10", "chrUn_gl0", "chrUn_gl000220", "*", "scaffold_" + "x" * 50, "y" * 70]
    lines = ["@HD\tVN:1.0", "@SQ\tSN:chr1\tLN:1000"]
    for k in range(30000):
        ops = []
        n_ops = int(rng.choice([1, 1, 1, 2, 3, 5, 8, 20, 40]))
        for i in range(n_ops):
            ops.append((int(rng.integers(1, 200)), "MIDNSHPX"[int(rng.integers(8))] if n_ops > 1 else "M"))
        if ops[0][1] == "N":
            ops[0] = (ops[0][0], "M")
        frag = sum(n for n, op in ops if op in "MISX")
        star_cigar = rng.random() < 0.05
        cigar = "*" if star_cigar else "".join("%d%s" % t for t in ops)
        if star_cigar:
            seq = "A" * int(rng.integers(1, 80))
        else:
            seq = "*" if (rng.random() < 0.1 or frag == 0) else "".join("ACGTN"[int(x)] for x in rng.integers(0, 5, frag))
        pos = 0 if rng.random() < 0.03 else int(rng.integers(1, 10 ** int(rng.integers(1, 10))))
        label = ["read%d" % k, "r", " lead", "a b", "-7", "123456789"][int(rng.integers(6))] if rng.random() < 0.3 else "q%d" % k
        opt = ["", "\tNM:i:1", "\tNM:i:1\tXS:A:+", "\tNM:i:1\tXS:A:+\tMD:Z:50"][int(rng.integers(4))]
        qual = "*" if rng.random() < 0.5 or seq == "*" else "I" * len(seq)
        lines.append("%s\t%d\t%s\t%d\t%d\t%s\t%s\t%d\t%d\t%s\t%s%s" % (label, int(rng.integers(0, 4096)), names[int(rng.integers(len(names)))], pos,
                                                                    int(rng.integers(0, 61)), cigar, ["*", "=", "chr2"][int(rng.integers(3))],
                                                                    int(rng.integers(0, 1000)), int(rng.integers(-500, 500)), seq, qual, opt))
    path = tmp_path / "shapes.sam"
    path.write_text("\n".join(lines) + "\n")
    want = ref_reg(path)
    assert want[0] == 0 and want[1].count(b"\n") > 25000                     # (a CIGAR that consumes no reference leaves a region without intervals: nothing is printed for it)
    for env in ({"GT_PARSE_THREADS": "1"}, {"GT_PARSE_THREADS": "5", "GT_PARSE_PIECE_BYTES": "777"}):
        got = dump(path, env)
        assert got[0] == 0 and strip_extras(got[1]) == want[1], env
    # labels as weights go through the general tokeniser: the same regions
    got_w = dump(path, {}, args=("-w", "5"))
    assert got_w[0] == 0 and strip_extras(got_w[1]) == want[1]
