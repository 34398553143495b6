#!/usr/bin/env python
"""CPU check of the closed-form zero-gradient Adam replay used by update_GMM (csrc/em.cu replay_explicit_steps /
replay_tail): explicit terms while beta1^s >= 1e-3 (1e-6 during the optimiser's first 2000 steps), then ONE geometric tail term, against the exact step-by-step
recursion in float64.  Prints the error relative to the replay's total movement."""
import math

lr, b1, b2, eps = 3e-3, 0.9, 0.999, 1e-8


def exact(p, m, v, first, count):
    for s in range(1, count + 1):
        t = first + s
        m, v = b1 * m, b2 * v
        p -= (lr / (1 - b1 ** t)) * m / (math.sqrt(v) / math.sqrt(1 - b2 ** t) + eps)
    return p


def closed(p, m, v, first, count):
    thr = 1e-3 if first >= 2000 else 1e-6            # young optimiser: 1 - b2^t still moves fast, keep more explicit terms
    S = min(count, max(1, math.ceil(math.log(thr) / math.log(b1))))
    a = math.sqrt(v)
    for s in range(1, S + 1):
        t = first + s
        p -= lr * b1 ** s / (1 - b1 ** t) * m / (a * math.sqrt(b2 ** s) / math.sqrt(1 - b2 ** t) + eps)
    if count > S:
        geo = b1 ** (S + 1) * (1 - b1 ** (count - S)) / (1 - b1)
        c = lr * geo / (1 - b1 ** (first + S + 1))
        ss = min(count, S + 1 + int(b1 / (1 - b1)))
        p -= c * m / (a * math.sqrt(b2 ** ss) / math.sqrt(1 - b2 ** (first + ss)) + eps)
    return p


if __name__ == "__main__":
    worst = 0.0
    for first in (0, 3, 50, 1000, 1999, 2000, 2500, 100000):
        for count in (1, 5, 66, 67, 70, 131, 450, 600, 3000):
            for m0, v0 in ((1e-3, 5e-7), (0.0145, 2.2e-5), (1e-6, 1e-12), (0.03, 1e-18)):
                pe, pc = exact(0.1, m0, v0, first, count), closed(0.1, m0, v0, first, count)
                worst = max(worst, abs(pe - pc) / max(abs(0.1 - pe), 1e-30))
    print("closed-form replay vs exact recursion: worst error / movement = %.2e" % worst)
    assert worst < 1e-6
