"""Kernel-bank sizing -- mirror of the reference's ``OS_CNN/OS_CNN_Structure_build.py`` (same function
names, argument names -- including the reference's spelling ``paramenter`` -- and return values).
Host-side integer arithmetic only; runs once at model construction."""


def get_Prime_number_in_a_range(start, end):
    """Values v in [start, end] with no divisor in [2, v).  Like the reference (file lines 3-13) this
    counts 1 as prime: the omni-scale bank always contains the 1-tap kernel."""
    return [v for v in range(start, end + 1) if not any(v % d == 0 for d in range(2, v))]


def get_out_channel_number(paramenter_layer, in_channel, prime_list):
    """Out channels per prime so that the layer holds about ``paramenter_layer`` weights (lines 16-18)."""
    return int(paramenter_layer / (in_channel * sum(prime_list)))


def generate_layer_parameter_list(start, end, paramenter_number_of_layer_list, in_channel=1):
    """List of layers, each a list of ``(in_ch, out_ch, kernel_size)``; one layer per budget, then a closing
    layer of the two kernels ``start`` and ``start + 1`` (lines 20-42)."""
    prime_list = get_Prime_number_in_a_range(start, end)
    if prime_list == []:
        print('start = ', start, 'which is larger than end = ', end)
    first_in_channel = in_channel
    layer_parameter_list = []
    for budget in paramenter_number_of_layer_list:
        out_channel = get_out_channel_number(budget, in_channel, prime_list)
        layer_parameter_list.append([(in_channel, out_channel, prime) for prime in prime_list])
        in_channel = len(prime_list) * out_channel
    closing_out = len(prime_list) * get_out_channel_number(paramenter_number_of_layer_list[0], first_in_channel, prime_list)
    layer_parameter_list.append([(in_channel, closing_out, start), (in_channel, closing_out, start + 1)])
    return layer_parameter_list
