from accel_rl_b200.buffers.batch import (batch_buffer, buffer_with_segs_view, buffer_length, view_segments,
                                         combine_distinct_buffers, count_buffer_size)
