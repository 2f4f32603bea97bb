"""Pins the oracle (and with it every parity claim of this repository) to the REAL reference binary.

  python tools/pin_reference.py --ref oracle/_ref/talc [--configs 1,3,5,k30small] [--workdir /tmp/talc_pin]

For each configuration: writes the synthetic inputs as the text files the reference reads (Jellyfish `dump -c`
text, optional junction dump, FASTA), runs the reference `talc ... -t 1 > /dev/null` (stdout flood
discarded, -t 1 for a deterministic log) and the oracle's own front end on the same files, and compares
`.fa` byte for byte, `.log` line for line (sorted), `.config.txt` modulo the output prefix, and the `.fa` / `.log`
digests with tests/golden/oracle_sha256.json.  Exit code 0 = the oracle is pinned.

The reference cannot be built in this image (SeqAn2 headers absent): tools/pin_reference.sh builds it
when SEQAN_INCLUDE points at a SeqAn2 checkout and then calls this script."""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SPECS = {  # name -> (config index, scale, reads, junctions, overrides, pin name or None)
    "1": (1, 1.0, 10000, False, {}, "config1"),
    "2": (2, 1.0, 131072, False, {}, "config2_batch"),
    "3": (3, 1.0, 32768, True, {}, "config3_batch"),
    "3small": (3, 0.004, 250, True, {}, "config3_small"),
    "4": (4, 1.0, 131072, False, {}, None),
    "5": (5, 1.0, 10000, False, {}, "config5"),
}


def sha(path, sort_lines=False):
    data = open(path, "rb").read() if os.path.exists(path) else b""
    if sort_lines:
        data = b"".join(sorted(data.splitlines(keepends=True)))
    return hashlib.sha256(data).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True, help="path to the reference `talc` binary")
    ap.add_argument("--configs", default="1,3small,5")
    ap.add_argument("--workdir", default="/tmp/talc_pin")
    args = ap.parse_args()
    from oracle import pyoracle as po
    from talc_b200 import synth
    po.build()
    pins_path = os.path.join(ROOT, "tests", "golden", "oracle_sha256.json")
    pins = json.load(open(pins_path)) if os.path.exists(pins_path) else {}
    os.makedirs(args.workdir, exist_ok=True)
    failed = []
    for name in args.configs.split(","):
        ci, scale, n, usej, over, pin = SPECS[name]
        cfg = synth.baseline_config(ci, scale)
        cfg.n_reads = n
        for k, v in over.items():
            setattr(cfg, k, v)
        d = os.path.join(args.workdir, "cfg" + name)
        os.makedirs(d, exist_ok=True)
        w = synth.make_workload(cfg)
        synth.write_dump(os.path.join(d, "sr.dump"), w.keys, w.counts, cfg.k)
        if usej:
            synth.write_dump(os.path.join(d, "j.dump"), w.jkeys, w.jcounts, cfg.k)
        synth.write_fasta(os.path.join(d, "reads.fa"), w.reads, w.read_off)
        common = ["reads.fa", "--SRCounts", "sr.dump", "-k", str(cfg.k)] + (["--junctions", "j.dump"] if usej else [])
        for prefix, exe, extra in (("ref", os.path.abspath(args.ref), []), ("orc", po.BIN, ["--oracle-table", "hash"])):
            for ext in (".log", ".fa"):
                if os.path.exists(os.path.join(d, prefix + ext)):
                    os.remove(os.path.join(d, prefix + ext))  # the log is opened in append mode (io.cpp:108)
            rc = subprocess.call([exe] + common + ["-o", prefix, "-t", "1"] + extra, cwd=d, stdout=subprocess.DEVNULL)
            if rc != 0:
                print("config %s: %s exited with %d" % (name, exe, rc))
        ok = True
        for ext, srt in ((".fa", False), (".log", True)):
            a, b = sha(os.path.join(d, "ref" + ext), srt), sha(os.path.join(d, "orc" + ext), srt)
            same = a == b
            print("config %-7s %-5s reference %s oracle %s  %s" % (name, ext, a[:16], b[:16], "same" if same else "DIFFERENT"))
            ok &= same
            if pin and pin in pins:
                key = "fa_sha256" if ext == ".fa" else "log_sha256"
                pinned = pins[pin][key] == a
                print("config %-7s %-5s committed pin %s  %s" % (name, ext, pins[pin][key][:16], "same" if pinned else "DIFFERENT"))
                ok &= pinned
        ca = open(os.path.join(d, "ref.config.txt")).read().replace("ref", "X") if os.path.exists(os.path.join(d, "ref.config.txt")) else None
        cb = open(os.path.join(d, "orc.config.txt")).read().replace("orc", "X")
        print("config %-7s .config.txt %s" % (name, "same" if ca == cb else "DIFFERENT"))
        ok &= ca == cb
        if not ok:
            failed.append(name)
    print("PINNED" if not failed else "NOT PINNED: configs %s differ" % failed)
    sys.exit(0 if not failed else 1)


if __name__ == "__main__":
    main()
