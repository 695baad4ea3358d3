"""Record types of the data layer (reference: src/data/basetypes.py:34-37): one depth frame with its annotations, and
a named sequence of frames with its crop configuration.  Field names and order are the reference's, so code that
builds or unpacks these tuples (importers, Dataset, the entry scripts) works unchanged."""
from collections import namedtuple

DepthFrame = namedtuple('DepthFrame', ['dpt', 'gtorig', 'gtcrop', 'T', 'gt3Dorig', 'gt3Dcrop', 'com', 'fileName',
                                       'subSeqName', 'side', 'extraData'])
NamedImgSequence = namedtuple('NamedImgSequence', ['name', 'data', 'config'])
