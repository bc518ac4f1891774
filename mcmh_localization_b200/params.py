"""Parameter keys of the reference node (app/params/amhmcl.yaml, read at
app/scripts/amcmh_localizer.py:18-58,84-85).  Same keys, same meaning."""

# defaults hard-coded in the node's rospy.get_param(name, default) calls (node:18-58, 84-85)
DEFAULT_PARAMS = dict(
    localization_mode="MHAMCL", initialized=False, init_particles=2000,
    alpha1=0.2, alpha2=0.2, alpha3=0.2, alpha4=0.2, alpha_slow=0.01, alpha_fast=0.1,
    kld_epsilon=0.025, kld_delta=0.99, kld_bin_size_xy=0.1, kld_bin_size_theta=0.17453292519943295,
    kld_z=2, min_particles=100, max_particles=5000,
    sigma_hit=0.2, max_range=10.0, z_hit=0.8, z_rand=0.2, step=1,
)

# the values shipped in app/params/amhmcl.yaml
YAML_PARAMS = dict(
    localization_mode="AMHAMCL", initialized=False, init_particles=1500,
    alpha1=0.002, alpha2=0.03, alpha3=0.08, alpha4=0.002,
    kld_epsilon=0.03, kld_z=2, kld_bin_size_xy=0.20, kld_bin_size_theta=0.1745, kld_delta=0.99,
    min_particles=100, max_particles=5000, alpha_slow=0.04, alpha_fast=0.6,
    sigma_hit=0.3, z_hit=0.75, z_rand=0.25, max_range=5.0, step=1,
)


def load_params(path=None, overrides=None):
    """DEFAULT_PARAMS overlaid with a rosparam-style YAML file (if given) and overrides."""
    p = dict(DEFAULT_PARAMS)
    if path is not None:
        import yaml
        with open(path) as f:
            p.update(yaml.safe_load(f) or {})
    if overrides:
        p.update(overrides)
    return p


def mode_flags(mode):
    """node:19-21: substring decoding of localization_mode."""
    return dict(use_mh="MH" in mode, use_adaptive="AMCL" in mode, assym="AMH" in mode)
