// Zsb200.cs -- P/Invoke declarations of libzsb200 (include/zsb200.h).  Source a maintainer of SixLabors.ZlibStream would add as
// src/ZlibStream/Native/Zsb200.cs; there is no .NET toolchain in the build image, so this file is not compiled here (the same ABI
// is exercised from Python, zlibstream_b200/_native.py, and by tests/test_abi.py).
// <copyright file="Zsb200.cs">Licensed under the Apache License, Version 2.0.</copyright>

using System;
using System.Runtime.InteropServices;

namespace SixLabors.ZlibStream.Native
{
    internal static unsafe class Zsb200
    {
        private const string Lib = "zsb200";   // libzsb200.so / zsb200.dll next to the assembly

        // flags (zsb200.h)
        public const uint InDevice = 0x1, OutDevice = 0x2, NoHeader = 0x4, NoTrailer = 0x8, OffsetsDevice = 0x10,
                          Continuation = 0x20, SingleStream = 0x40, SyncFlush = 0x80;

        // Deflate.Intrinsics.cs:295-307: which UpdateHash this build of the reference runs
        public const int HashCrc32C = 0, HashMul = 1;

        [DllImport(Lib)] public static extern int zsb_ctx_create(int device, out IntPtr ctx);            // new ZLibStream(options)
        [DllImport(Lib)] public static extern void zsb_ctx_destroy(IntPtr ctx);                          // ZLibStream.Dispose
        [DllImport(Lib)] public static extern IntPtr zsb_last_error(IntPtr ctx);                         // ZLibStream.Message
        [DllImport(Lib)] public static extern int zsb_ctx_set_stream(IntPtr ctx, IntPtr cudaStream);
        [DllImport(Lib)] public static extern UIntPtr zsb_deflate_bound(UIntPtr n, UIntPtr segSize);

        // foreach segment: ZlibOutputStream.Write(segment) with FlushMode.FullFlush (or SyncFlush), then Dispose/Finish;
        // with SingleStream: one Write with FlushMode.NoFlush, then Dispose
        [DllImport(Lib)] public static extern int zsb_deflate_segments(
            IntPtr ctx, byte* input, UIntPtr n, UIntPtr segSize, int level, int strategy, int hashVariant,
            ulong firstSegmentIndex, byte* halo, uint flags,
            byte* output, UIntPtr outCap, out UIntPtr outLen, ulong* segEndOffsets, out uint adler);

        // the loop over ZlibInputStream.Read
        [DllImport(Lib)] public static extern int zsb_inflate_segments(
            IntPtr ctx, byte* input, UIntPtr n, ulong* segEndOffsets, UIntPtr nseg, UIntPtr segSize, uint flags,
            byte* output, UIntPtr outCap, out UIntPtr outLen, out uint adler);

        [DllImport(Lib)] public static extern int zsb_ctx_last_inflate_carry(IntPtr ctx, out uint firstSegmentCarry);
        [DllImport(Lib)] public static extern int zsb_inflate_fill_carry(IntPtr ctx, byte prevByte, out uint adler);

        // segment index of a stream that has no side-car (InflateSync's marker, Inflate.cs:386-448)
        [DllImport(Lib)] public static extern int zsb_build_segment_index(
            IntPtr ctx, byte* input, UIntPtr n, UIntPtr segSize, uint flags,
            ulong* segEndOffsets, UIntPtr cap, out UIntPtr nseg, out ulong uncompressedSize);

        // the PNG encoder above the reference (ImageSharp): filter + deflate + CRC + framing on the device
        [DllImport(Lib)] public static extern UIntPtr zsb_png_bound(uint width, uint height, int colorType, UIntPtr segSize, UIntPtr idatBytes);
        [DllImport(Lib)] public static extern int zsb_png_encode(
            IntPtr ctx, byte* pixels, uint width, uint height, int colorType, int filter, int level, int strategy,
            UIntPtr segSize, UIntPtr idatBytes, uint flags, byte* output, UIntPtr outCap, out UIntPtr outLen);

        [DllImport(Lib)] public static extern int zsb_adler32(IntPtr ctx, uint seed, byte* p, UIntPtr n, uint flags, out uint result);
        [DllImport(Lib)] public static extern uint zsb_adler32_combine(uint adler1, uint adler2, ulong len2);

        /// <summary>CompressionState other than ZOK / ZSTREAMEND becomes the exception the reference throws (ThrowHelper.cs:21-23).</summary>
        public static void Check(IntPtr ctx, int rc, bool isCompressing)
        {
            if (rc != (int)CompressionState.ZOK && rc != (int)CompressionState.ZSTREAMEND)
            {
                string msg = Marshal.PtrToStringAnsi(zsb_last_error(ctx));
                ThrowHelper.ThrowCompressionException(isCompressing, msg);
            }
        }
    }
}
