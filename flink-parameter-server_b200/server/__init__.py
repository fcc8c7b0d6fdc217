from .logics import (LockPSLogicA, LockPSLogicB, LooseSimplePSLogic, LooseSimplePSLogicWithClose,
                     RangePSLogicWithClose, SimplePSLogic, SimplePSLogicWithClose)
