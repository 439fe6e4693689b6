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
t1 = time.time()
print("rc", r.returncode, "wall %.2fs" % (t1 - t0))
# where the wall time goes outside the log's own span: process start-up before the first line,
# teardown (device memory release) after the last
import datetime
stamps = []
for l in r.stderr.splitlines():
    try:
        stamps.append(datetime.datetime.strptime(l[:26], "%Y/%m/%d %H:%M:%S.%f").timestamp())
    except ValueError:
        pass
if stamps:
    print("start-up %.2fs  logged span %.2fs  teardown %.2fs" % (stamps[0] - t0, stamps[-1] - stamps[0], t1 - stamps[-1]))
for l in r.stderr.splitlines():
    print(l[:160])
print(r.stdout[:300].replace("\n", " | "))
