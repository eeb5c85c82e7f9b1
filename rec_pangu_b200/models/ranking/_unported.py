"""Reference ranking classes that are outside the north-star hot path (AFN, AOANet,
CCPM are not named by BASELINE.json's north_star; MaskNet and LR are built: masknet.py, lr.py).  The names exist so that
`from rec_pangu.models.ranking import ...` lines in the reference's examples import unchanged; constructing one
fails loudly instead of silently running a non-B200 path."""


def _unported(name):
    class _Unported:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                f'{name} is not part of the B200 hot-path scope (SURVEY.md §8); use one of WDL, DeepFM, NFM, '
                f'FiBiNet, AFM, AutoInt, FM, xDeepFM, DCN, MaskNet, LR')
    _Unported.__name__ = name
    return _Unported


AFN = _unported('AFN')
AOANet = _unported('AOANet')
CCPM = _unported('CCPM')
