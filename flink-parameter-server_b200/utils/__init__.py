from .eof import EOF, EOFHandler, flatMapWithEOF, with_eof
from .input_source import EventWithTimestamp, InputSource
from .sleep_blocker import block
from . import metrics, model_io, tracing
