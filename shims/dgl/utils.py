def expand_as_pair(feat, g=None):
    return feat, feat
