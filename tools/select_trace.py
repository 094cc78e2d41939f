#!/usr/bin/env python
"""Prints the phase trace of the ordered-selection kernel (OFB_SELECT_TRACE=1) for one 1080p and one 4K frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["OFB_SELECT_TRACE"] = "1"
import ofb200, synth
ctx = ofb200.Context(0)
for (h, w, k) in [(1080, 1920, 1000), (2160, 3840, 5000), (720, 1280, 500)]:
    img = synth.texture(h, w, 0)
    for rep in range(2):
        p = ofb200.goodFeaturesToTrack(img, k, 0.01, 10, blockSize=7, ctx=ctx)
    print(h, w, k, len(p), flush=True)
