"""Latency of one dependent random 32-byte sector as a function of the footprint it is drawn from (is the probe of a
multi-GB successor table slowed by address translation?).  usage: python tools/probe_latency.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from talc_b200 import api
p = api.TalcParams()
api.lib().talc_params_default(p, 21)
t = api.Talc(p, 0)
res = {}
for mb in (64, 256, 1024, 2048, 4096, 8192, 16384):
    row = {}
    for wps in (1, 8):
        g, ns = t.bench_random_sectors(mb << 20, True, wps)
        row["hop_ns_%d_warps_per_sm" % wps] = round(ns, 1)
        row["gbs_%d_warps_per_sm" % wps] = round(g, 1)
    res["%d MiB" % mb] = row
print(json.dumps(res))
