class Evaluator:  # placeholder
    def __init__(self, *a, **k):
        raise NotImplementedError("stub")


class PygGraphPropPredDataset:  # placeholder
    def __init__(self, *a, **k):
        raise NotImplementedError("stub")
