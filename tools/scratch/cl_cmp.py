import sys, numpy as np
a = np.load(sys.argv[1], allow_pickle=True); b = np.load(sys.argv[2], allow_pickle=True)
for name in a["__order"]:
    name = str(name)
    x, y = a[name], b[name]
    bad = ~((x == y) | (np.isnan(x) & np.isnan(y)))
    if bad.any():
        Bn, C, L = x.shape
        print(name, x.shape, "bad frac", bad.mean())
        print("  bad per batch item:", bad.reshape(Bn, -1).mean(1).round(2))
        print("  bad per 32-channel block:", bad.transpose(1, 0, 2).reshape(C // 32, -1).mean(1).round(2))
        nb = min(L // 128, 16) or 1
        print("  bad per 128-row block (item 0):", bad[0].reshape(C, -1)[:, : nb * 128].reshape(C, nb, -1).transpose(1, 0, 2).reshape(nb, -1).mean(1).round(2))
        print("  bad per 128-row block (item 1):", bad[1].reshape(C, -1)[:, : nb * 128].reshape(C, nb, -1).transpose(1, 0, 2).reshape(nb, -1).mean(1).round(2))
        break
else:
    print("all equal")
