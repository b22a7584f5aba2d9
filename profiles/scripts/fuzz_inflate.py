import os, random, struct, subprocess, sys, zlib
TOOL = sys.argv[1]
random.seed(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
tmp = "/dev/shm/fz_%d" % os.getpid()
os.makedirs(tmp, exist_ok=True)
def run(data, bufsize=None):
    p = os.path.join(tmp, "x.gz")
    open(p, "wb").write(data)
    a = [TOOL, p] + ([str(bufsize)] if bufsize else [])
    r = subprocess.run(a, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    if r.returncode not in (0, 1):
        raise SystemExit("crash rc=%d stderr=%s" % (r.returncode, r.stderr[-2000:].decode(errors="replace")))
    return r.returncode, r.stdout
def gz_member(raw, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=15, memlevel=8, flags=0):
    c = zlib.compressobj(level, zlib.DEFLATED, -wbits, memlevel, strategy)
    body = c.compress(raw) + c.flush()
    hdr = b"\x1f\x8b\x08" + bytes([flags]) + b"\0\0\0\0\0\xff"
    if flags & 4: hdr += struct.pack("<H", 7) + b"AB\x03\x00xyz"
    if flags & 8: hdr += b"file name.txt\0"
    if flags & 16: hdr += b"a comment\0"
    if flags & 2: hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
    return hdr + body + struct.pack("<II", zlib.crc32(raw) & 0xFFFFFFFF, len(raw) & 0xFFFFFFFF)
def ref_prefix(data):
    """what zlib hands out for a (possibly truncated) series of members"""
    out = b""
    rest = data
    first = True
    while rest:
        if rest[:2] != b"\x1f\x8b":
            break
        d = zlib.decompressobj(31)
        try:
            out += d.decompress(rest)
        except zlib.error:
            return None
        if not d.eof:
            break
        rest = d.unused_data
        first = False
    return out
def make_raw(kind, n):
    if kind == "rand": return os.urandom(n)
    if kind == "text": return b"".join(b"chr%d\t%d\t%d\tr%d\t0\t%s\n" % (random.randrange(1, 23), random.randrange(10**8), random.randrange(10**8), i, random.choice([b"+", b"-"])) for i in range(n // 30 + 1))[:n]
    if kind == "rle": return bytes([random.randrange(4)]) * (n // 2) + os.urandom(3) * (n // 6 + 1)
    if kind == "mixed":
        parts = []
        while sum(map(len, parts)) < n:
            k = random.randrange(4)
            parts.append([os.urandom(random.randrange(1, 70000)), b"ab" * random.randrange(1, 40000), make_raw("text", random.randrange(1, 90000)), bytes(random.randrange(1, 300000))][k])
        return b"".join(parts)[:n]
n_cases = 0
sizes = [0, 1, 2, 100, 32767, 32768, 32769, 65536, 1 << 20, (1 << 20) + 1, (1 << 20) - 258, 3_000_017]
for kind in ("rand", "text", "rle", "mixed"):
    for n in sizes:
        raw = make_raw(kind, n)
        for level, strategy, wbits in ((0, 0, 15), (1, 0, 15), (6, 0, 15), (9, 0, 15), (6, zlib.Z_FIXED, 15), (6, zlib.Z_HUFFMAN_ONLY, 15), (6, zlib.Z_RLE, 15), (9, zlib.Z_FILTERED, 9), (4, 0, 12)):
            if n > (1 << 20) + 1 and level in (9,) and kind == "rand": continue
            flags = random.choice([0, 0, 8, 4 | 8 | 16, 2, 4 | 2])
            data = gz_member(raw, level, strategy, wbits, random.choice([1, 8, 9]), flags)
            rc, out = run(data, random.choice([None, 1, 7, 4096, 65536, 1 << 20, (1 << 20) + 3]) if n < 200000 else None)
            assert rc == 0 and out == raw, ("whole", kind, n, level, strategy, wbits, flags, len(out))
            n_cases += 1
            # truncations
            for _ in range(3 if n < 200000 else 1):
                cut = random.randrange(0, len(data))
                want = ref_prefix(data[:cut])
                rc, out = run(data[:cut])
                assert want is not None and out == want, ("cut", kind, n, level, strategy, wbits, cut, len(out), len(want))
                n_cases += 1
            # damage: must not crash; whatever comes out before the fault is a prefix of... nothing is promised but no crash and rc 0/1
            for _ in range(2 if n < 200000 else 1):
                if len(data) < 30: break
                bad = bytearray(data)
                for _ in range(random.randrange(1, 4)):
                    bad[random.randrange(10, len(bad))] ^= 1 << random.randrange(8)
                rc, out = run(bytes(bad))
                # (rc 0 with other bytes is possible: damage that makes the data run off the end of the file reads as a truncated file)
                n_cases += 1
# several members, empty members, trailing garbage
for trial in range(30):
    raws = [make_raw(random.choice(["rand", "text", "rle", "mixed"]), random.choice([0, 1, 1000, 70000, 1 << 20, 1500000])) for _ in range(random.randrange(1, 5))]
    data = b"".join(gz_member(r, random.choice([0, 1, 6, 9]), flags=random.choice([0, 8, 4])) for r in raws)
    rc, out = run(data)
    assert rc == 0 and out == b"".join(raws), ("members", trial)
    rc, out = run(data + b"\0\0\0trailing garbage" * random.randrange(1, 5))
    assert out == b"".join(raws), ("garbage", trial)
    cut = random.randrange(0, len(data) + 1)
    rc, out = run(data[:cut])
    assert out == ref_prefix(data[:cut]), ("members cut", trial, cut)
    n_cases += 3
print("ok", n_cases, "cases")
