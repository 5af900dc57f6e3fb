"""Import stub so that /root/reference/lib_yolo/utils.py and detect.py can be imported (plotting is out of scope)."""
