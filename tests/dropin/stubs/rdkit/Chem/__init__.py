class _Mol:
    def __init__(self, smiles):
        self.smiles = smiles


def MolFromSmiles(smiles):
    if not isinstance(smiles, str) or smiles == "":
        raise ValueError("bad smiles")
    return _Mol(smiles)
