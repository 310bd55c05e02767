def spherical_harmonics(degrees_to_use, dirs, coeffs, masks=None):
    if dirs.is_cuda:
        from bilateral_driving_b200.render import spherical_harmonics as product

        return product(degrees_to_use, dirs, coeffs, masks)
    from oracle.sh_ref import spherical_harmonics as oracle_sh

    out = oracle_sh(degrees_to_use, dirs, coeffs)
    return out if masks is None else out * masks[..., None]
