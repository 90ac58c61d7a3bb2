"""Decode /root/reference/data/nsclc.rda (the reference's own example dataset, R/nnmf.R:110-115) into
tests/golden/nsclc.npz without R. Run in the build container only (the GPU box has no /root/reference).

Format: bzip2 -> "RDX2\\nX\\n" XDR serialisation v2 -> pairlist(tag 'nsclc' -> REALSXP with attributes dim, dimnames).
We only need the 20000 big-endian doubles and dim = c(200, 100).
"""
import bz2
import struct
import sys

import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/data/nsclc.rda"
raw = bz2.decompress(open(SRC, "rb").read())
assert raw[:5] == b"RDX2\n" and raw[5:7] == b"X\n", raw[:8]
# locate the REALSXP header: flags word whose low byte is 14 (REALSXP) followed by length 20000
needle = struct.pack(">i", 20000)
pos = -1
i = 7
while True:
    i = raw.find(needle, i)
    if i < 0:
        break
    flags = struct.unpack(">i", raw[i - 4:i])[0]
    if flags & 0xFF == 14:
        pos = i + 4
        break
    i += 1
assert pos > 0, "REALSXP(20000) not found"
vals = np.frombuffer(raw[pos:pos + 8 * 20000], dtype=">f8").astype(np.float64)
A = vals.reshape((200, 100), order="F")
assert np.isfinite(A).all()
print("nsclc", A.shape, A.min(), A.max(), A.mean())
assert abs(A.min() - 2.59627) < 1e-4 and abs(A.max() - 14.10538) < 1e-4 and abs(A.mean() - 7.0201492) < 1e-6
np.savez_compressed("tests/golden/nsclc.npz", A=A)
