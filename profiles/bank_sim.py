import numpy as np
rng = np.random.default_rng(0)
T = 20000
def rows(n):
    cells = np.where(rng.random((n,4)) < 0.3, 0, rng.integers(1, 12, (n,4)))
    return cells[:,0] | cells[:,1]<<4 | cells[:,2]<<8 | cells[:,3]<<12
def wavefronts(idx):
    # idx: (T,32) word indices; wavefronts = max over banks of distinct addresses
    out = np.zeros(len(idx))
    for t in range(len(idx)):
        u = np.unique(idx[t])
        out[t] = np.bincount(u & 31, minlength=32).max()
    return out.mean()
r = rows(T*32).reshape(T,32)
print("plain          ", wavefronts(r))
print("uniform ref    ", wavefronts(rng.integers(0,57344,(T,32))))
for name, f in [
    (">>5", lambda x: x ^ ((x>>5)&31)),
    (">>8", lambda x: x ^ ((x>>8)&31)),
    (">>4", lambda x: x ^ ((x>>4)&31) & ~0 ),
    (">>5^>>10", lambda x: x ^ ((x>>5)&31) ^ ((x>>10)&31)),
    (">>6", lambda x: x ^ ((x>>6)&31)),
    (">>7", lambda x: x ^ ((x>>7)&31)),
    (">>9", lambda x: x ^ ((x>>9)&31)),
    (">>11", lambda x: x ^ ((x>>11)&31)),
    (">>3 (bits3..7 -> 0..4)", lambda x: x ^ ((x>>3)&31 & 0b11110)),
]:
    print(f"{name:15s}", wavefronts(f(r)))
