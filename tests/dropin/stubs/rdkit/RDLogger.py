def DisableLog(spec):
    return None
