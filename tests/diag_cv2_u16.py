"""Diagnostic: is cv2.GaussianBlur(u16) on this host the exact integer formula?"""
import subprocess, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from oracle import restate
print(cv2.__version__, cv2.getNumThreads(), cv2.useOptimized())
print(subprocess.run("lscpu | grep -E 'Model name|^CPU\\(s\\)|Flags' | cut -c1-1500", shell=True, capture_output=True, text=True).stdout)
rng = np.random.default_rng(0)
for (w, h) in [(7, 9), (64, 64), (320, 256)]:
    for hi in (256, 4096, 16384, 32768, 40000, 65536):
        img = rng.integers(0, hi, (h, w)).astype(np.uint16)
        ex = restate.blur5(img)
        cv2.setUseOptimized(True)
        a = cv2.GaussianBlur(img, (5, 5), 0)
        cv2.setUseOptimized(False)
        b = cv2.GaussianBlur(img, (5, 5), 0)
        cv2.setUseOptimized(True)
        print((w, h), hi, 'optimized!=exact', int((a != ex).sum()), 'unoptimized!=exact', int((b != ex).sum()),
              'maxdiff', int(np.abs(a.astype(int) - ex.astype(int)).max()))
# where do the mismatches sit?
img = rng.integers(0, 65536, (40, 64)).astype(np.uint16)
a = cv2.GaussianBlur(img, (5, 5), 0); ex = restate.blur5(img)
ys, xs = np.nonzero(a != ex)
print('mismatch coords (first 20):', list(zip(ys[:20].tolist(), xs[:20].tolist())))
for y, x in list(zip(ys, xs))[:5]:
    print('at', (y, x), 'cv2', int(a[y, x]), 'exact', int(ex[y, x]), 'window:\n', img[max(y-2,0):y+3, max(x-2,0):x+3])
# smooth realistic u16 frame
from cylinder_pose_estimation_b200 import synth
f = synth.render_u16(320, 256, seed=0, n=9, pitch=14.0)
print('realistic u16 frame mismatches:', int((cv2.GaussianBlur(f, (5, 5), 0) != restate.blur5(f)).sum()), 'max pixel', int(f.max()))
u8 = rng.integers(0, 256, (64, 64)).astype(np.uint8)
print('u8 mismatches:', int((cv2.GaussianBlur(u8, (5, 5), 0) != restate.blur5(u8)).sum()))
