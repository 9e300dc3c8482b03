"""Oracle-backed stand-in for the engine's State-cache calls (TEST INFRASTRUCTURE: tests/ and the CPU columns of the benches only).

`rpmplv2_b200.drgbt.generate_gbur` and similar drivers are written against store_distance / store_extend_spines / distance; this
class answers them with the CPU oracle (computeDistance / extendGenSpine restated in oracle/rvs_oracle.hpp), i.e. what the
reference computes one call at a time."""
import numpy as np

from rpmplv2_b200 import robots


class OracleSpace:
    """store_distance / store_extend_spines / distance on top of the CPU oracle (tests only)."""

    def __init__(self, o):
        self.o, self.slots = o, {}

    def store_reserve(self, capacity):
        pass

    def store_distance(self, q, slots):
        d, p, near = self.o.distance(q, robots.DIST_FULL)
        for k, s in enumerate(slots):
            self.slots[int(s)] = (q[k].copy(), float(d[k]), p[k].copy(), near[k].copy())
        return d, p

    def store_extend_spines(self, root_slots, q_targets, sp, gen2=False):
        uniq = sorted(set(int(s) for s in root_slots))
        idx = np.array([uniq.index(int(s)) for s in root_slots], np.int32)
        return self.o.extend_spines_layers(np.stack([self.slots[s][0] for s in uniq]), np.stack([self.slots[s][2] for s in uniq]),
                                           np.array([self.slots[s][1] for s in uniq], np.float32), np.stack([self.slots[s][3] for s in uniq]),
                                           idx, q_targets, sp, gen2)

    def distance(self, q, mode=robots.DIST_PROFILE):
        return self.o.distance(q, mode)
