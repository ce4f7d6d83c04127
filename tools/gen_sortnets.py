"""Verifies (0-1 principle) and prints the comparator networks used by clip_mean's group update
(csrc/rollout_consensus.cu, SortNet<N>) and checks the half-cleaner selection property."""
import itertools
import random

NETS = {
    2: [(0, 1)],
    3: [(0, 1), (1, 2), (0, 1)],
    4: [(0, 1), (2, 3), (0, 2), (1, 3), (1, 2)],
    5: [(0, 1), (3, 4), (2, 4), (2, 3), (1, 4), (0, 3), (0, 2), (1, 3), (1, 2)],
    6: [(1, 2), (4, 5), (0, 2), (3, 5), (0, 1), (3, 4), (2, 5), (0, 3), (1, 4), (2, 4), (1, 3), (2, 3)],
    7: [(1, 2), (3, 4), (5, 6), (0, 2), (3, 5), (4, 6), (0, 1), (4, 5), (2, 6), (0, 4), (1, 5), (0, 3), (2, 5), (1, 3), (2, 4), (2, 3)],
    8: [(0, 2), (1, 3), (4, 6), (5, 7), (0, 4), (1, 5), (2, 6), (3, 7), (0, 1), (2, 3), (4, 5), (6, 7), (2, 4), (3, 5), (1, 4),
        (3, 6), (1, 2), (3, 4), (5, 6)],
}
for n, net in NETS.items():
    for bits in itertools.product([0, 1], repeat=n):
        a = list(bits)
        for i, j in net:
            if a[i] > a[j]:
                a[i], a[j] = a[j], a[i]
        assert a == sorted(a), (n, bits)
    print(f"template <> struct SortNet<{n}> {{ template <bool ASC> static __device__ __forceinline__ void run(float (&a)[{n}]) {{ "
          + " ".join(f"cswap<ASC>(a[{i}], a[{j}]);" for i, j in net) + " } };")
for K in range(1, 9):
    for _ in range(2000):
        A, B = sorted(random.choices(range(20), k=K)), sorted(random.choices(range(20), k=8))
        assert sorted(min(A[i], B[K - 1 - i]) for i in range(K)) == sorted(A + B)[:K]
        Ad = sorted(A, reverse=True)
        assert sorted((max(Ad[i], B[8 - K + i]) for i in range(K)), reverse=True) == sorted(A + B, reverse=True)[:K]
print("// all networks sort; half-cleaner selection verified")
