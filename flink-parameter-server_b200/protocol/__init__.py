from .combination import (Combinable, CombinationLogic, CountLogic, TimerLogic, all_of, any_of)
from .messages import PSToWorker, Pull, PullAnswer, Push, WorkerToPS
from .senders import (CombinationPSSender, CombinationWorkerSender, CountClientSender, CountPSSender,
                      MultiplePSReceiver, MultipleWorkerReceiver, PSReceiver, PSSender,
                      SimplePSReceiver, SimplePSSender, SimpleWorkerReceiver, SimpleWorkerSender,
                      TimerClientSender, TimerPSSender, WorkerReceiver, WorkerSender)
