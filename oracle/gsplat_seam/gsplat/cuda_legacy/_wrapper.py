def num_sh_bases(degree: int) -> int:
    assert degree <= 4
    return (degree + 1) ** 2
