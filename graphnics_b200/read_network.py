"""Reader for the vascular-network text files of Secomb's group, drop-in for
data/secomb_tumours/read_network.py:14-57.  `get_graph` takes a URL (needs network
access and `requests`) or a local path; `parse_network` takes the text itself."""
import os
import re

import networkx as nx

from .fenics_graph import copy_from_nx_graph

__all__ = ["get_graph", "parse_network"]


def parse_network(text):
    """table rows: ... v1 v2 diameter <col> x1 y1 z1 x2 y2 z2 (counted from the right: [-10] [-9] [-8] | [-6:-3] [-3:]);
    the first two lines are header, the last two are trailer (read_network.py:28-46)"""
    rows = text.split("\n")[2:-2]
    H = nx.DiGraph()
    for row in rows:
        vals = re.split(r"\s{1,}", " " + row)
        v1, v2 = int(vals[-10]), int(vals[-9])
        H.add_edge(v1, v2)
        H.edges()[(v1, v2)]["radius"] = float(vals[-8]) / 2
        H.nodes()[v1]["pos"] = [float(c) for c in vals[-6:-3]]
        H.nodes()[v2]["pos"] = [float(c) for c in vals[-3:]]
    return copy_from_nx_graph(nx.convert_node_labels_to_integers(H))


def get_graph(url):
    if os.path.exists(url):
        with open(url) as fh:
            return parse_network(fh.read())
    import requests
    print(f"Downoading data from {url} \n")
    print("Please visit https://physiology.arizona.edu/people/secomb/network for information on how to cite "
          "the source for this data\n  ")
    return parse_network(requests.get(url, allow_redirects=True).text)
