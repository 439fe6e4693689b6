#!/usr/bin/env python
"""Developer tool: end-to-end `camus` CLI run on a synthetic configuration, phase times from its log."""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camus_b200 import build, synth
cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
s, c = synth.make_config(cid, *(int(x) for x in sys.argv[2:4])) if len(sys.argv) > 3 else synth.make_config(cid)
d = tempfile.mkdtemp()
ext = "nex" if s.fmt == "nexus" else "nwk"
open(f"{d}/c.nwk", "w").write(s.constraint + "\n")
open(f"{d}/g.{ext}", "w").write(s.genes)
cmd = [build.CLI_BIN, "-o", f"{d}/out", "-q", str(c["qmode"]), "-t", str(c["threshold"]), "-sm", c["mode"], "-f", s.fmt]
if c["as_set"] and c["mode"] != "sym":
    cmd.append("-asSet")
if c["min_support"]:
    cmd += ["-s", str(c["min_support"])]
cmd += [f"{d}/c.nwk", f"{d}/g.{ext}"]
t0 = time.time()
r = subprocess.run(cmd, capture_output=True, text=True)
print("rc", r.returncode, "wall %.2fs" % (time.time() - t0))
for l in r.stderr.splitlines():
    print(l[:160])
print(r.stdout[:300].replace("\n", " | "))
