python tools/exp_two_streams.py 1 512
KEEPB200_SMS=74 python tools/exp_two_streams.py 2 512
KEEPB200_SMS=74 python tools/exp_two_streams.py 2 256
KEEPB200_SMS=148 python tools/exp_two_streams.py 2 512
KEEPB200_SMS=112 python tools/exp_two_streams.py 2 512
python tools/exp_two_streams.py 1 512
