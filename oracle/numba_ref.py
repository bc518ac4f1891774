"""CPU baseline leg that runs the REFERENCE'S OWN numba code (test / measurement infrastructure, like the rest of
oracle/; never imported by the product).

The reference is Python + numba and is neither vendored nor copied: when its tree is present (this build
container: /root/reference, or $MCL_REFERENCE) app/scripts/parallel_utils.py is imported read-only through
sys.path, exactly as oracle/gen_golden.py does, and the filter step of the node (amcmh_localizer.py: odom_callback
:379-408, lidar_callback :294-338) is replayed with those functions -- BASELINE.md section 3's protocol.  On a box
without the tree (the GPU boxes) `available()` says why and the caller falls back to the C port.
"""
import os
import sys

import numpy as np

from . import node_glue as ng

REF = os.environ.get("MCL_REFERENCE", "/root/reference")
_pu = None
_why = None


def available():
    """(module or None, reason).  Imports the reference's parallel_utils on first use."""
    global _pu, _why
    if _pu is not None or _why is not None:
        return _pu, _why
    path = os.path.join(REF, "app", "scripts")
    if not os.path.isfile(os.path.join(path, "parallel_utils.py")):
        _why = "reference tree not present on this box (%s)" % path
        return None, _why
    try:
        import numba  # noqa: F401
    except Exception as e:  # pragma: no cover
        _why = "numba not importable: %s" % e
        return None, _why
    sys.path.insert(0, path)
    try:
        import parallel_utils as pu
    except Exception as e:  # pragma: no cover
        _why = "reference import failed: %s" % e
        return None, _why
    finally:
        sys.path.pop(0)
    _pu = pu
    return _pu, None


class NumbaReferenceFilter(ng.ReferenceFilter):
    """ReferenceFilter whose per-particle arithmetic is the reference's own njit functions (its own MT19937
    draws: this class is for TIMING; parity goes through the injected-draw restatements)."""

    def __init__(self, mp, params, particles, mode="MHMCL", threads=None):
        super().__init__(mp, params, particles, mode)
        pu, why = available()
        if pu is None:
            raise RuntimeError(why)
        import numba
        self.pu = pu
        self.threads = int(threads or len(os.sched_getaffinity(0)))
        numba.set_num_threads(min(self.threads, numba.config.NUMBA_NUM_THREADS))
        self.threads = numba.get_num_threads()
        self.layer = None

    def move_particles(self, odom, normals=None, seed=0, step=0):
        current = np.asarray(odom, dtype=np.float64)
        if self.last_odom is not None:
            self.delta = ng.compute_motion(self.last_odom, current)
            mp = self.mp
            prop = self.pu.apply_motion_model_parallel(self.particles, self.delta, self.alpha, mp["map_data"], mp["resolution"],
                                                       mp["origin_np"][0], mp["origin_np"][1], mp["width"], mp["height"])
            self.particles_prev = self.particles.copy()                    # node:404-405
            self.particles = prop.copy()
        self.last_odom = current

    def likelihood(self, particles, scan, angles):
        mp, p = self.mp, self.p
        return self.pu.compute_likelihoods(scan, angles, particles, mp["distance_map"], mp["resolution"], mp["origin_np"],
                                           mp["width"], mp["height"], p["sigma_hit"], p["z_hit"], p["z_rand"], p["max_range"],
                                           p["step"])

    def update(self, scan, angles, uniforms=None, seed=0, step=0):
        scores_pre = self.likelihood(self.particles_prev, scan, angles)      # node:254-259
        weights_pre = ng.convert_scores(scores_pre)
        scores_post = self.likelihood(self.particles, scan, angles)          # node:263-268
        weights_post = ng.convert_scores(scores_post)
        if self.use_mh:
            self.particles, weights = self.pu.mh_resampling(self.particles_prev, self.particles, weights_post, weights_pre)
        else:
            weights = weights_post
        self.weights = weights
        return weights

    def resample(self, r=None):
        n = len(self.particles)
        self.particles, _ = self.pu.low_variance_resample_numba(self.particles, self.weights, n)   # node:488-492

    def describe(self):
        import numba
        return "numba %s (%s threading layer, %d threads), numpy %s" % (numba.__version__, numba.threading_layer(), self.threads,
                                                                        np.__version__)
