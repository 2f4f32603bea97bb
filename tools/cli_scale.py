"""BASELINE config 4 through the PRODUCT: the `talc` command line with --gpus 1 / 2 / ... on one set of reads of the
full-size workload (table from the text dump, replicated by the library's NCCL broadcast, batches dealt round-robin).
Checks that every run writes the same .fa / .log and reports wall seconds (process start -> files closed).
usage: python tools/cli_scale.py [reads=1000000] [gpus=1,2]"""
import hashlib, json, os, shutil, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from talc_b200 import api, build, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
gpus = [x for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2").split(",")]  # "2p" = two devices, --replicate peer
cfg = synth.baseline_config(2, 1.0)
dev = "cuda:0"
tr = synth.make_transcriptome(cfg, dev)
keys, counts, jk, jc = synth.make_counts(cfg, tr, dev)
reads, roff = synth.make_reads(cfg, tr, n_reads, dev, seed_offset=4)
d = tempfile.mkdtemp(prefix="talc_cli_scale_")
api.write_dump(os.path.join(d, "sr.dump"), keys.cpu().numpy().astype(np.uint64), counts.cpu().numpy(), cfg.k)
synth.write_fasta(os.path.join(d, "reads.fa"), reads, roff)
bases = int(roff[-1])
del tr, keys, counts, reads, roff
torch.cuda.empty_cache()
cli = build.build_cli()
res = {"reads": n_reads, "bases": bases, "runs": {}}
digests = set()
for g in gpus:
    t0 = time.time()
    peer = g.endswith("p")
    ng = int(g.rstrip("p"))
    pr = subprocess.run([cli, "reads.fa", "--SRCounts", "sr.dump", "-k", str(cfg.k), "-o", "g%s" % g, "--gpus", str(ng), "-t", "16"] +
                        (["--replicate", "peer"] if peer else []), cwd=d, capture_output=True, text=True)
    rc = pr.returncode
    secs = time.time() - t0
    phases = [l for l in pr.stdout.splitlines() if "seconds:" in l or "replicated" in l]
    fa = hashlib.sha256(open(os.path.join(d, "g%s.fa" % g), "rb").read()).hexdigest()
    lg = hashlib.sha256(open(os.path.join(d, "g%s.log" % g), "rb").read()).hexdigest() if os.path.exists(os.path.join(d, "g%s.log" % g)) else ""
    digests.add((fa, lg))
    res["runs"]["gpus_%s" % g] = {"rc": rc, "seconds": round(secs, 2), "mbp_per_s": round(bases / 1e6 / secs, 1), "fa_sha256": fa[:16], "phases": phases}
res["identical_outputs"] = len(digests) == 1
shutil.rmtree(d, ignore_errors=True)
print(json.dumps(res))
