import rdkit


class _BitVect:
    def __init__(self, bits):
        self.bits = bits


def GetMorganFingerprintAsBitVect(mol, radius, nBits=2048):
    return _BitVect(rdkit.fake_morgan_bits(mol.smiles, nBits))
