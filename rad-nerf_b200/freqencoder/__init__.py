from .freq import FreqEncoder
