// ZlibInputStream.Gpu.cs -- the read side of the shim: replaces ReadCore / Finish of ZlibInputStream.cs:133-186,230-266 when
// libzsb200 is present.  Mirrors zlibstream_b200/streams.py; not compiled here (no .NET).
// <copyright file="ZlibInputStream.Gpu.cs">Licensed under the Apache License, Version 2.0.</copyright>

using System;
using System.IO;
using SixLabors.ZlibStream.Native;

namespace SixLabors.ZlibStream
{
    /// <summary>
    /// Reads ahead a bounded batch of compressed bytes from <c>BaseStream</c> (the reference reads 8192 bytes ahead,
    /// ZlibInputStream.cs:150-161), cuts it at a full-flush marker, lets the device find and decode the segments in it and
    /// serves <c>Read</c> from the decoded batch.  A chance occurrence of 00 00 FF FF inside compressed data makes the index
    /// call fail (its last piece does not end at a flush point); the batch is then retried with more input.
    /// </summary>
    public sealed unsafe partial class ZlibInputStream
    {
        private static readonly byte[] Marker = { 0x00, 0x00, 0xFF, 0xFF };
        private IntPtr ctx;
        private int segSize = 65536, batchSegments = 4096;
        private byte[] ahead = new byte[1 << 20];
        private int aheadLen;
        private byte[] decoded = Array.Empty<byte>();
        private int decodedLen, decodedPos;
        private bool first = true, eofIn, done;
        private uint adler = 1;

        private int GpuReadCore(Span<byte> buffer)
        {
            while (this.decodedPos >= this.decodedLen && !this.done) { this.NextBatch(); }
            int n = Math.Min(buffer.Length, this.decodedLen - this.decodedPos);
            this.decoded.AsSpan(this.decodedPos, n).CopyTo(buffer);
            this.decodedPos += n;
            return n;
        }

        private void Pull(int want)
        {
            while (this.aheadLen < want && !this.eofIn)
            {
                if (this.ahead.Length < want) { Array.Resize(ref this.ahead, Math.Max(want, this.ahead.Length * 2)); }
                int r = this.BaseStream.Read(this.ahead, this.aheadLen, Math.Max(8192, want - this.aheadLen));
                if (r == 0) { this.eofIn = true; break; }
                this.aheadLen += r;
            }
        }

        private void NextBatch()
        {
            int want = Math.Max(1 << 16, this.batchSegments * (this.segSize / 2));
            while (true)
            {
                this.Pull(want);
                int hdr = this.first ? 2 : 0;
                int cut;
                if (this.eofIn)
                {
                    bool storedFinal = this.aheadLen >= 9 && this.ahead[this.aheadLen - 9] == 1 && this.ahead[this.aheadLen - 6] == 0xFF;
                    int tl = storedFinal ? 9 : 6;                             // final block + Adler-32 (level 0: empty stored block)
                    if (this.aheadLen <= hdr + tl) { this.CheckTrailer(hdr); return; }
                    cut = this.aheadLen - tl;
                }
                else
                {
                    cut = this.ahead.AsSpan(hdr, this.aheadLen - 9 - hdr).LastIndexOf(Marker);
                    if (cut < 0) { want *= 2; continue; }
                    cut += hdr + 4;
                }

                int maxSeg = cut / 64 + 16;
                ulong[] ends = new ulong[maxSeg];
                UIntPtr nseg;
                ulong total;
                uint flags = (this.first ? 0u : Zsb200.NoHeader) | Zsb200.NoTrailer;
                int rc;
                fixed (byte* pin = this.ahead)
                fixed (ulong* pe = ends)
                {
                    rc = Zsb200.zsb_build_segment_index(this.ctx, pin, (UIntPtr)cut, (UIntPtr)this.segSize, flags, pe, (UIntPtr)maxSeg, out nseg, out total);
                    if (rc != 0 && !this.eofIn) { want *= 2; continue; }      // the cut was not a flush point: read further
                    Zsb200.Check(this.ctx, rc, isCompressing: false);
                    if (this.decoded.Length < (int)total) { this.decoded = new byte[(int)total]; }
                    fixed (byte* pout = this.decoded)
                    {
                        rc = Zsb200.zsb_inflate_segments(this.ctx, pin, (UIntPtr)cut, pe, nseg, (UIntPtr)this.segSize, flags,
                            pout, (UIntPtr)total, out UIntPtr outLen, out uint a);
                        Zsb200.Check(this.ctx, rc, isCompressing: false);
                        this.adler = Zsb200.zsb_adler32_combine(this.adler, a, (ulong)outLen);
                        this.decodedLen = (int)outLen;
                        this.decodedPos = 0;
                    }
                }

                Buffer.BlockCopy(this.ahead, cut, this.ahead, 0, this.aheadLen - cut);
                this.aheadLen -= cut;
                this.first = false;
                return;
            }
        }

        /// <summary>Inflate.cs:286-342: the final block, then the big-endian Adler-32 ("incorrect data check").</summary>
        private void CheckTrailer(int at)
        {
            ReadOnlySpan<byte> t = this.ahead.AsSpan(at, this.aheadLen - at);
            int skip = t.Length >= 2 && t[0] == 3 && t[1] == 0 ? 2 : (t.Length >= 5 && t[0] == 1 && t[3] == 0xFF ? 5 : -1);
            if (skip < 0) { ThrowHelper.ThrowCompressionException(false, "invalid block type"); }
            if (t.Length < skip + 4) { ThrowHelper.ThrowCompressionException(false, "buffer error"); }
            uint want = ((uint)t[skip] << 24) | ((uint)t[skip + 1] << 16) | ((uint)t[skip + 2] << 8) | t[skip + 3];
            if (want != this.adler) { ThrowHelper.ThrowCompressionException(false, "incorrect data check"); }
            this.done = true;
        }
    }
}
