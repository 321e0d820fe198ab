def copy_src(src, out):
    return ('copy_src', src, out)


def u_mul_e(u, e, out):
    return ('u_mul_e', u, e, out)


def sum(msg, out):   # noqa: A001 - the name DGL uses
    return ('sum', msg, out)
