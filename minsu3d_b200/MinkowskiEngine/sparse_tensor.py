"""SparseTensor + coordinate manager (T1, T2) on top of the libb2s coordinate hash.

Semantics follow SURVEY.md appendix A items 1-6 and 15 (restated from MinkowskiEngine 0.5.4):
coordinates are int32 [M, 4] with the batch index in column 0; duplicates keep the first row;
a stride-2 map keeps coordinates in input units (multiples of the tensor stride); kernel maps are
cached per (in stride, out stride, kernel size) for the lifetime of the manager, which is created
once per SparseTensor(features, coordinates) construction (backbone.py:38, general_model.py:191).
"""
import torch

from .. import ops


class CoordinateMapKey:
    """Identifies a coordinate map inside one manager (tensor stride + string id)."""

    __slots__ = ("stride", "string_id")

    def __init__(self, stride, string_id=""):
        self.stride = int(stride)
        self.string_id = string_id

    def get_tensor_stride(self):
        return [self.stride] * 3

    def get_key(self):
        return (self.get_tensor_stride(), self.string_id)

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and (self.stride, self.string_id) == (other.stride, other.string_id)

    def __hash__(self):
        return hash((self.stride, self.string_id))

    def __repr__(self):
        return "CoordinateMapKey(stride=%d)" % self.stride


class _CoordMap:
    __slots__ = ("coords", "table", "stride")

    def __init__(self, coords, table, stride):
        self.coords = coords
        self.table = table
        self.stride = stride

    @property
    def size(self):
        return self.coords.size(0)


class KernelMap:
    """Output-stationary neighbour table + lazily built canonical pair lists."""

    def __init__(self, nbr, n_in, n_out, exact_pairs=None, tile_mask=None):
        self.nbr = nbr
        self.tile_mask = tile_mask  # per-128-row active-offset masks (ops.kernel_map), None = let the kernel scan
        self.n_in = n_in
        self.n_out = n_out
        self.K = nbr.size(1)
        self._pairs = None
        self._exact = exact_pairs
        self._sorted = None

    # Large 3^3 maps run the tcgen05 convolution over mask-sorted tiles (ops.tile_order): in first-occurrence row
    # order every 128-row tile sees all 27 offsets, sorted tiles see ~13 on ScanNet-shaped scenes.  The sort costs
    # about one convolution and is shared by the ~16 forward / data-gradient launches on the map.
    SORTED_MIN_ROWS = 32768

    def sorted_tables(self, c_in, c_out, algo=None):
        """(nbr_sorted, tile_mask_sorted, row_perm) when the sorted schedule applies to this map, else Nones."""
        nbr, tile_mask, out_rows = self.table_for(c_in, c_out, algo)
        return (nbr, tile_mask, out_rows) if out_rows is not None else (None, None, None)

    def table_for(self, c_in, c_out, algo=None):
        """(nbr, tile_mask, out_rows) to hand to ops.conv_table for a c_in -> c_out product on this map."""
        algo = ops.get_conv_algo() if algo is None else algo
        if (self.K == 27 and self.n_out >= self.SORTED_MIN_ROWS and self.n_in == self.n_out
                and algo != ops.ALGO_SIMT and ops.conv_tc_shape_ok(self.K, c_in, c_out)):
            if self._sorted is None:
                self._sorted = ops.tile_order(self.nbr)
            row_perm, nbr_sorted, tile_mask = self._sorted
            return nbr_sorted, tile_mask, row_perm
        return self.nbr, self.tile_mask, None

    def pairs(self):
        """(pair_in, pair_out, k_offsets, max_pairs): sorted by (kernel offset, output row)."""
        if self._pairs is None:
            pin, pout, koff, _ = ops.pairs_from_nbr(self.nbr)
            max_pairs = self._exact if self._exact is not None else self.n_out * self.K
            self._pairs = (pin, pout, koff, max_pairs)
        return self._pairs


class CoordinateManager:
    def __init__(self, D=3, device=None):
        self.D = D
        self.device = device
        self._maps = {}
        self._kmaps = {}
        self._size_hints = None  # {tensor stride: rows} reported by the voxeliser (see SparseTensor(level_sizes=...))

    # -- maps ---------------------------------------------------------------------------------
    def insert_and_map(self, coordinates, tensor_stride=1, assume_unique=False):
        """Register coordinates; returns (key, unique_index i32, inverse_mapping i32)."""
        table, unique_idx, inverse, out_coords = ops.coord_unique(coordinates, quant=1, assume_unique=assume_unique)
        key = CoordinateMapKey(tensor_stride)
        self._maps[key] = _CoordMap(out_coords, table, int(tensor_stride))
        self.device = coordinates.device
        return key, unique_idx, inverse

    def get_coordinates(self, key):
        return self._maps[key].coords

    def size(self, key):
        return self._maps[key].size

    # a MinkUNet strides down several times: when the first stride-2 map of a large tensor is requested the
    # whole coordinate pyramid is built speculatively on the device with a single host read of all row counts
    # (each level is at most half the previous one, unused levels are cheap); small tensors build one level.
    PYRAMID_MIN_ROWS = 20000
    PYRAMID_LEVELS = 6

    def stride_key(self, in_key, stride=2):
        """Coordinate map of tensor stride in*stride: unique(floor(c / new) * new), first-occurrence order."""
        new_stride = in_key.stride * int(stride)
        out_key = CoordinateMapKey(new_stride)
        if out_key not in self._maps:
            src = self._maps[in_key]
            levels = 1
            if int(stride) == 2 and src.size >= self.PYRAMID_MIN_ROWS:
                levels = self.PYRAMID_LEVELS
            if levels > 1:
                hints = None
                if self._size_hints is not None:
                    want = [in_key.stride * 2 ** (i + 1) for i in range(levels)]
                    if all(w in self._size_hints for w in want):
                        hints = [self._size_hints[w] for w in want]
                for s, table, oc in ops.coord_pyramid(src.coords, in_key.stride, levels, size_hints=hints):
                    key = CoordinateMapKey(s)
                    if key not in self._maps:
                        self._maps[key] = _CoordMap(oc, table, s)
            else:  # small tensors: one level at a time, one host read each (size hints are used by the pyramid only)
                table, _, _, out_coords = ops.coord_unique(src.coords, quant=new_stride)
                self._maps[out_key] = _CoordMap(out_coords, table, new_stride)
        return out_key

    def existing_key(self, stride):
        key = CoordinateMapKey(stride)
        if key not in self._maps:
            raise ValueError("no coordinate map with tensor stride %d in this manager" % stride)
        return key

    # -- kernel maps --------------------------------------------------------------------------
    def kernel_map(self, in_key, out_key, kernel_size):
        """nbr[o, kidx] = input row at out_coords[o] + offset(kidx) * in_stride (appendix A.4-5)."""
        ck = (in_key, out_key, int(kernel_size))
        km = self._kmaps.get(ck)
        if km is None:
            src, dst = self._maps[in_key], self._maps[out_key]
            nbr, tile_mask = ops.kernel_map(dst.coords, src.table, int(kernel_size), src.stride, with_tile_mask=True)
            exact = src.size if (out_key.stride != in_key.stride) else None  # 2^3/s2: one parent per input row
            km = KernelMap(nbr, src.size, dst.size, exact_pairs=exact, tile_mask=tile_mask)
            self._kmaps[ck] = km
        return km


class SparseTensor:
    """Sparse tensor = features [M, C] float32 + a coordinate map key inside a manager."""

    def __init__(self, features, coordinates=None, tensor_stride=1, coordinate_map_key=None,
                 coordinate_manager=None, quantization_mode=None, device=None, coordinates_unique=False,
                 level_sizes=None, **_unused):
        """coordinates_unique / level_sizes are extensions (not in MinkowskiEngine): the caller states that the
        rows are already unique (they come out of sparse_quantize) and, optionally, how many rows the strided
        maps {tensor stride: rows} will have (the voxeliser knows).  Both remove host reads of device counts, so
        the host keeps enqueuing while the GPU is still busy; the statements are validated on the device and a
        wrong one raises at the next host read (ops.run_deferred_checks)."""
        if device is not None:
            features = features.to(device)
        self.quantization_mode = quantization_mode
        self.inverse_mapping = None
        self.unique_index = None
        if coordinate_map_key is None:
            if coordinates is None:
                raise ValueError("either coordinates or coordinate_map_key must be given")
            if isinstance(tensor_stride, (list, tuple)):
                tensor_stride = tensor_stride[0]
            if not torch.is_tensor(coordinates):
                coordinates = torch.as_tensor(coordinates)
            if coordinates.is_floating_point():
                coordinates = torch.floor(coordinates)
            coordinates = coordinates.to(device=features.device, dtype=torch.int32).contiguous()
            if not features.is_cuda:
                raise ValueError("SparseTensor needs CUDA features (libb2s has no CPU path)")
            if coordinates.size(0) != features.size(0):
                raise ValueError("coordinates and features have different numbers of rows")
            if coordinate_manager is None:
                coordinate_manager = CoordinateManager(D=coordinates.size(1) - 1, device=features.device)
            if level_sizes is not None:
                coordinate_manager._size_hints = {int(k): int(v) for k, v in dict(level_sizes).items()}
            coordinate_map_key, unique_idx, inverse = coordinate_manager.insert_and_map(
                coordinates, tensor_stride, assume_unique=bool(coordinates_unique))
            self.inverse_mapping = inverse
            self.unique_index = unique_idx
            if unique_idx.numel() != coordinates.size(0):
                # duplicates: keep the first row's features (RANDOM_SUBSAMPLE, appendix A.2)
                features = features[unique_idx.long()]
        elif coordinate_manager is None:
            raise ValueError("coordinate_map_key needs a coordinate_manager")
        self._F = features
        self.coordinate_map_key = coordinate_map_key
        self.coordinate_manager = coordinate_manager

    # -- accessors ----------------------------------------------------------------------------
    @property
    def F(self):
        return self._F

    @property
    def features(self):
        return self._F

    @features.setter
    def features(self, value):
        self._F = value

    @property
    def C(self):
        return self.coordinate_manager.get_coordinates(self.coordinate_map_key)

    @property
    def coordinates(self):
        return self.C

    @property
    def tensor_stride(self):
        return self.coordinate_map_key.get_tensor_stride()

    @property
    def device(self):
        return self._F.device

    @property
    def dtype(self):
        return self._F.dtype

    @property
    def D(self):
        return self.coordinate_manager.D

    @property
    def shape(self):
        return self._F.shape

    def size(self, *a):
        return self._F.size(*a)

    def __len__(self):
        return self._F.size(0)

    def _like(self, features):
        return SparseTensor(features, coordinate_map_key=self.coordinate_map_key,
                            coordinate_manager=self.coordinate_manager)

    def _check_same_map(self, other):
        if (self.coordinate_manager is not other.coordinate_manager
                or self.coordinate_map_key != other.coordinate_map_key):
            raise ValueError("sparse tensors live on different coordinate maps")

    # -- arithmetic on identical coordinate maps (common.py:48) ---------------------------------
    def __add__(self, other):
        if isinstance(other, SparseTensor):
            self._check_same_map(other)
            return self._like(self._F + other._F)
        return self._like(self._F + other)

    def __iadd__(self, other):
        if isinstance(other, SparseTensor):
            self._check_same_map(other)
            self._F = self._F + other._F
        else:
            self._F = self._F + other
        return self

    def __sub__(self, other):
        if isinstance(other, SparseTensor):
            self._check_same_map(other)
            return self._like(self._F - other._F)
        return self._like(self._F - other)

    def __mul__(self, other):
        if isinstance(other, SparseTensor):
            self._check_same_map(other)
            return self._like(self._F * other._F)
        return self._like(self._F * other)

    def __repr__(self):
        return "SparseTensor(rows=%d, channels=%d, tensor_stride=%s, device=%s)" % (
            self._F.size(0), self._F.size(1), self.tensor_stride, self._F.device)
