#!/usr/bin/env python3
"""Dev-container check (needs /root/reference): the oracle's rebuilt MQ state table equals the
94 rows of internal/entropy/mqc.go:21-116, and melExp / UVLC prefix rows match ht.go / ht_luts.go.
Not run by pytest on the GPU box (the reference is absent there)."""
import ctypes as C
import os
import re
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import oracle_lib as O  # noqa: E402

src = open("/root/reference/internal/entropy/mqc.go").read()
rows = re.findall(r"\{0x([0-9A-Fa-f]+), (\d), (\d+), (\d+)\},\s*// (\d+)", src)
assert len(rows) == 94, len(rows)
L = O.lib()
L.orc_mq_tables_init()
qe = (C.c_uint32 * 94).in_dll(L, "orc_mq_qe")
nm = (C.c_uint8 * 94).in_dll(L, "orc_mq_nmps")
nl = (C.c_uint8 * 94).in_dll(L, "orc_mq_nlps")
for q, mps, nmps, nlps, idx in rows:
    i = int(idx)
    assert (qe[i], i & 1, nm[i], nl[i]) == (int(q, 16), int(mps), int(nmps), int(nlps)), i
print("MQ state table: 94/94 rows identical to mqc.go:21-116")
