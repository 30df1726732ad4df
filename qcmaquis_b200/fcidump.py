import numpy as np, sys
def make_fcidump(path, n, nelec, seed=20240607):
    rng = np.random.default_rng(seed)
    h = -np.diag(np.linspace(1, 0.2, n)); a = rng.standard_normal((n, n)); h = h + 0.05 * (a + a.T) / 2
    Lv = []
    for P in range(3 * n):
        a = rng.standard_normal((n, n)); Lv.append(0.3 * (a + a.T) / 2 / np.sqrt(n))
    Lv = np.array(Lv)
    eri = np.einsum('pij,pkl->ijkl', Lv, Lv)
    with open(path, 'w') as f:
        f.write("&FCI NORB=%d,NELEC=%d,MS2=0,\n ORBSYM=%s\n ISYM=1,\n&END\n" % (n, nelec, ",".join(["1"] * n) + ","))
        for i in range(n):
            for j in range(i + 1):
                for k in range(i + 1):
                    for l in range(k + 1):
                        if i * (i + 1) // 2 + j >= k * (k + 1) // 2 + l:
                            f.write("%23.16e %d %d %d %d\n" % (eri[i, j, k, l], i + 1, j + 1, k + 1, l + 1))
        for i in range(n):
            for j in range(i + 1):
                f.write("%23.16e %d %d 0 0\n" % (h[i, j], i + 1, j + 1))
        f.write("%23.16e 0 0 0 0\n" % 0.5)
if __name__ == "__main__":
    make_fcidump(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
