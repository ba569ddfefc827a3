def ConvertToNumpyArray(fp, array):
    """RDKit resizes the destination in place (the script passes np.zeros((0,), dtype=np.int8))."""
    array.resize(fp.bits.shape[0], refcheck=False)
    array[:] = fp.bits
