from shim_backend import qratio_percent


def QRatio(s1, s2, *, processor=None, score_cutoff=None):
    return qratio_percent(s1, s2)
