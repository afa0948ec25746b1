"""std::mt19937 + libstdc++ std::uniform_real_distribution<double>, so the reference's seeded
randomized tests (align_test.cpp:444-601, seeds 12345 / 6789 / 9999) can be replayed exactly."""


class MT19937:
    def __init__(self, seed):
        self.mt = [0] * 624
        self.idx = 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.idx = 0

    def next_u32(self):
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def canonical(self):
        """std::generate_canonical<double, 53>(mt19937): two 32-bit draws."""
        lo = self.next_u32()
        hi = self.next_u32()
        r = (lo + hi * 4294967296.0) / 18446744073709551616.0
        if r >= 1.0:
            r = 1.0 - 2.0 ** -53
        return r

    def uniform(self, a, b):
        return self.canonical() * (b - a) + a
