"""Configuration of the staytime VideoDnn multi-task model (reference: staytime/config.py).

Pure data: the feature-slot ids the model is wired on, the sequence slots, the stay-time histogram
(400 bins of 0.5 s from -19.0, i.e. `bin_list`), expert / task counts and task names."""

_SLOT_TEXT = """
1568 1570 1571 1574 1575 1576 1577 1578 1579 1581 1582 1583 1585 1587 1589 1591 1592 1593 1594 1595 1599 1601
1611 1612 1614 1616 1623 1636 1736 1737 1738 1739 1740 1741 1743 1744 1749 2039 2040 2041 2042 2043 2044
2050 2051 2052 2123 2125 2127 2128 2130 2131 2135 2139 2142 2144 2147 2149 2151 2152 2154 2156 2544
2597 3051 3365 3369 3376 3370 1745 2045 1632 1735 2153 2047 2244 2046 2150 2247 1625 1624 2148 2159
2146 2242 2260 2155 2259 2615 4500 4386
"""


class Configure(object):
    SLOTS = _SLOT_TEXT.split()                              # 91 sparse slots (staytime/config.py:4-14)
    SEQ_SLOTS = ["2125", "2128", "2130"]                    # behaviour sequences: videoid, authorid, l1 cate
    multiclass_num = 400
    bin_list = [-19.0 + 0.5 * i for i in range(400)]        # -19.0 ... 180.5
    num_experts = 3
    num_tasks = 3
    task_names = ["staytime_pred", "shortplay_pred", "longplay_pred"]
    model_name = "video_id_rank_staytime_mtl_ppnet_v7"


Config = Configure()
