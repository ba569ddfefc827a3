import rdkit


class _Rxn:
    def __init__(self, smarts):
        self.smarts = smarts


def ReactionFromSmarts(smarts):
    return _Rxn(smarts)


def CreateDifferenceFingerprintForReaction(rxn):
    return [int(v) for v in rdkit.fake_difference_counts(rxn.smarts)]
