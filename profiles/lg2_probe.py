import importlib, struct, sys, os
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("camera-aware-neural-networks-for-few-view-depth-estimation_b200")
def bits(x): return struct.unpack("<I", struct.pack("<f", x))[0]
for lo, hi in ((1e-7, 1100.0), (1e-6, 1000.0), (2.0**-16, 1100.0), (0.05, 20.0), (0.25, 4.0), (0.5, 2.0)):
    tol = 2.0 ** -17
    while pkg.selftest(2, bits(lo), bits(hi), tol * 2 ** -0.25) == 0:
        tol *= 2 ** -0.25
    print(f"lg2.approx.ftz max abs error over [{lo:g}, {hi:g}] <= {tol:.3e}")
