class _Stub:
    def __init__(self, *a, **k):
        raise RuntimeError("FrEIA is not installed; this is an import-only stub")


InputNode = ConditionNode = Node = OutputNode = GraphINN = ReversibleGraphNet = _Stub
