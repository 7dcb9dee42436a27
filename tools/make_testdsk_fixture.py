#!/usr/bin/env python
"""Extracts the known-answer vectors of gatb-core's DSK counting test (test/unit/src/kmer/TestDSK.cpp:147-240, `DSK_check1`)
into tests/golden/testdsk/check1.json: the literal test sequences and the (kmerSize, abundance-min, expected number of solid
k-mers) rows.  Run in the build container (reads /root/reference, which does not exist on the GPU box)."""
import json
import os
import re

SRC = "/root/reference/thirdparty/gatb-core/gatb-core/test/unit/src/kmer/TestDSK.cpp"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "testdsk", "check1.json")

txt = open(SRC).read()
body = txt[txt.index("void DSK_check1 ()"):txt.index("void DSK_check2_aux ()")]
s1 = re.search(r'const char\* s1 = "([ACGT]+)"', body).group(1)
arr = body[body.index("const char* seqs4[] = {"):body.index("} ;")]
seqs4 = ["".join(re.findall(r'"([ACGT]+)"', chunk)) for chunk in arr.split(",\n") if '"' in chunk]
sets = {"seqs1": [s1], "seqs2": [s1, s1], "seqs3": [s1, s1, s1], "seqs4": seqs4}
rows = [{"seqs": m.group(1), "k": int(m.group(2)), "abundance_min": int(m.group(3)), "nb_solid": int(m.group(4))}
        for m in re.finditer(r"DSK_check1_aux \((seqs\d), ARRAY_SIZE\(seqs\d\), (\d+), (\d+), (\d+)\);", body)]
assert len(seqs4) == 3 and len(rows) == 33, (len(seqs4), len(rows))
os.makedirs(os.path.dirname(OUT), exist_ok=True)
json.dump({"source": "gatb-core test/unit/src/kmer/TestDSK.cpp:147-240 (DSK_check1)", "sequences": sets, "rows": rows}, open(OUT, "w"), indent=1)
print(OUT, len(rows), "rows", [len(s) for s in seqs4])
