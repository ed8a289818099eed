"""Reader.get_variables_interpolated at the reader's own precision on the host build of the device sources: bit-equal to the
unmodified reference's StructuredReader (tests/readercases.py)."""
import readercases as rc
from hostengine import HostEngine


def test_reader_output_equals_the_reference_reader_bit_for_bit():
    assert rc.check(HostEngine()) == 27
